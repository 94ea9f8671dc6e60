#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for m in 0 1 2; do
NVO_ADAM_MODE=$m timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/c16_bench_m$m.json 2> gpurun_out/c16_bench_m$m.err
python -c "
import json
d=json.load(open('gpurun_out/c16_bench_m$m.json')); print('mode $m', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
NVO_ADAM_MODE=2 timeout 300 python tools/timeline.py --tag r2s7_adam2 --pose off > gpurun_out/timeline_r2s7.log 2>&1
