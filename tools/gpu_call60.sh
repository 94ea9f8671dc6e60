#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for e in default 0 1 2 default; do
if [ $e = default ]; then unset NVO_EARLY_FIELDS_OPT; else export NVO_EARLY_FIELDS_OPT=$e; fi
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c60_bench_e$e.json 2> gpurun_out/c60_bench_e$e.err
python -c "
import json
d=json.load(open('gpurun_out/c60_bench_e$e.json')); print('bench early=$e', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
unset NVO_EARLY_FIELDS_OPT
for lp in -1; do
NVO_LEVEL_PRIORITY=$lp timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c60_bench_lp.json 2> gpurun_out/c60_bench_lp.err
python -c "
import json
d=json.load(open('gpurun_out/c60_bench_lp.json')); print('bench level priority $lp', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
