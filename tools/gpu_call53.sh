#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for pr in 1 0 1 0; do
NVO_GRID_FWD_PAIR=$pr timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg > gpurun_out/c53_bench_pair$pr.json 2> gpurun_out/c53_bench_pair$pr.err
python -c "
import json
d=json.load(open('gpurun_out/c53_bench_pair$pr.json')); print('bench grid pair $pr', d['value'], d['ms_per_step'], d['e2e']['value'], [round(o['launch_us'],1) for o in d['roofline']['others'][:1]])"
done
