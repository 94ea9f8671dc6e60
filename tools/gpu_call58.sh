#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for v in base adam8 adam6 base adam8 adam6; do
cp build/libnvo_$v.so nerf-vo_b200/libnvo_b200.so
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg > gpurun_out/c58_bench_$v.json 2> gpurun_out/c58_bench_$v.err
python -c "
import json
d=json.load(open('gpurun_out/c58_bench_$v.json')); print('bench $v', d['value'], d['ms_per_step'], d['e2e']['value'], [round(o['launch_us'],1) for o in d['roofline']['others'][1:2]])"
done
