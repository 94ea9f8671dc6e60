#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c59_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c59_pytest.log
tail -3 gpurun_out/c59_pytest.log
for h in 1 0 1 0; do
NVO_FIELD_SCALE_HISTORY=$h timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c59_bench_h$h.json 2> gpurun_out/c59_bench_h$h.err
python -c "
import json
d=json.load(open('gpurun_out/c59_bench_h$h.json')); print('bench history $h', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'], d['final_loss'])"
done
