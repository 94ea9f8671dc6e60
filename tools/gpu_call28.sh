#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c28_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c28_pytest.log
tail -4 gpurun_out/c28_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/c28_bench.json 2> gpurun_out/c28_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c28_bench.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"
timeout 300 python tools/timeline.py --tag r2s11 --pose off > gpurun_out/timeline_r2s11.log 2>&1
python tools/show_timeline.py gpurun_out/timeline_r2s11.csv 0 | cut -c1-110
