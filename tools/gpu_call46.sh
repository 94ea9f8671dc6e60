#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_full_size_parity.py tests/test_trainer.py -m gpu -q -x 2>&1 | tail -3
for md in 3 2 3 2; do
NVO_ADAM_MODE=$md timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c46_bench_m$md.json 2> gpurun_out/c46_bench_m$md.err
python -c "
import json
d=json.load(open('gpurun_out/c46_bench_m$md.json')); print('bench adam mode $md', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"
done
