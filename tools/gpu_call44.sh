#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c44_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c44_pytest.log
tail -4 gpurun_out/c44_pytest.log
for pr in -1 0; do
NVO_OPT_PRIORITY=$pr timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c44_bench_p$pr.json 2> gpurun_out/c44_bench_p$pr.err
python -c "
import json
d=json.load(open('gpurun_out/c44_bench_p$pr.json')); print('bench opt priority $pr', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"
done
NVO_ADAM_MODE=2 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c44_bench_m2.json 2> gpurun_out/c44_bench_m2.err
python -c "
import json
d=json.load(open('gpurun_out/c44_bench_m2.json')); print('bench adam mode 2', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"
timeout 300 python tools/timeline.py --tag c44 --pose off > gpurun_out/timeline_c44.log 2>&1; tail -1 gpurun_out/timeline_c44.log
