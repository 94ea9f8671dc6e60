#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pose_opt.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/c14_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c14_pytest.log
tail -4 gpurun_out/c14_pytest.log
timeout 300 python tools/timeline.py --tag r2s6_pose > gpurun_out/timeline_r2s6_pose.log 2>&1
tail -2 gpurun_out/timeline_r2s6_pose.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/c14_bench_pose.json 2> gpurun_out/c14_bench_pose.err
python -c "
import json
for f in ('c14_bench_pose',):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d.get('reference_schedule',{}).get('value'), d['e2e']['value'], d['gpu_launches_per_step'])"
tail -3 gpurun_out/c14_bench_pose.err
