#!/bin/bash
# usage: tools/gpu_ab_env.sh VAR v1 v2 ...  -- same-box A/B of an environment switch: the default bench (config 2, 200 steps) once per value, twice over
VAR=$1; shift
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
env $VAR=$v timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/ab_${VAR}_$v.json 2> gpurun_out/ab_${VAR}_$v.err
python -c "
import json
d=json.load(open('gpurun_out/ab_${VAR}_$v.json')); print('$VAR=$v', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
done
done
