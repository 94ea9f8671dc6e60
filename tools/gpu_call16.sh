#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for m in 0 40960 57344 73728; do
NVO_ADAM_SMEM=$m timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/c16_bench_s$m.json 2> gpurun_out/c16_bench_s$m.err
python -c "
import json
d=json.load(open('gpurun_out/c16_bench_s$m.json')); print('smem $m', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
NVO_EARLY_FIELDS_OPT=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/c16_bench_noearly.json 2> gpurun_out/c16_bench_noearly.err
python -c "
import json
d=json.load(open('gpurun_out/c16_bench_noearly.json')); print('noearly', d['value'], d['ms_per_step'], d['e2e']['value'])"
