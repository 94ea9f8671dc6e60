#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c25_pytest.log 2>&1
echo "rc=$? wall=${SECONDS}s" >> gpurun_out/c25_pytest.log
tail -4 gpurun_out/c25_pytest.log
SECONDS=0
timeout 600 python bench.py > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
echo "bench rc=$? wall=${SECONDS}s"
python -c "
import json
d=json.load(open('gpurun_out/c25_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'], d['cpu_baseline']['value'], d.get('reference_schedule',{}).get('value'), d['pose_opt']['value'])"
for c in 4 5; do
SECONDS=0
timeout 600 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline --no-schedule-leg > gpurun_out/c25_bench_config$c.json 2> gpurun_out/c25_bench_config$c.err
echo "config$c rc=$? wall=${SECONDS}s"
python -c "
import json
d=json.load(open('gpurun_out/c25_bench_config$c.json')); print('config$c', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))"
tail -2 gpurun_out/c25_bench_config$c.err
done
SECONDS=0
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c25_bench_reference.json 2> gpurun_out/c25_bench_reference.err
echo "reference rc=$? wall=${SECONDS}s"; cut -c1-600 gpurun_out/c25_bench_reference.json
