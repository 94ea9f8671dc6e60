#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c42_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c42_pytest.log
tail -4 gpurun_out/c42_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c42_bench.json 2> gpurun_out/c42_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c42_bench.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"
timeout 300 python tools/timeline.py --tag c42 --pose off > gpurun_out/timeline_c42.log 2>&1; tail -1 gpurun_out/timeline_c42.log
