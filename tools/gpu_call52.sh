#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for pr in 1 0 1 0; do
NVO_PROP_FWD_PAIR=$pr timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c52_bench_pair$pr.json 2> gpurun_out/c52_bench_pair$pr.err
python -c "
import json
d=json.load(open('gpurun_out/c52_bench_pair$pr.json')); print('bench pair $pr', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
NVO_PROP_FWD_PAIR=0 timeout 300 python tools/timeline.py --tag c52 --pose off > gpurun_out/timeline_c52.log 2>&1; python tools/show_timeline.py gpurun_out/timeline_c52.csv 20 | grep prop_fwd
NVO_PROP_FWD_PAIR=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
