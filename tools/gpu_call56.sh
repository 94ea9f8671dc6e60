#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c56_bench$i.json 2> gpurun_out/c56_bench$i.err
python -c "
import json
d=json.load(open('gpurun_out/c56_bench$i.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
