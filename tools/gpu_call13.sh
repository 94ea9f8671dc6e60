#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pose_opt.py -m gpu -q -x > gpurun_out/c13_pose.log 2>&1
echo "rc=$?" >> gpurun_out/c13_pose.log
tail -30 gpurun_out/c13_pose.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_pose_opt.py > gpurun_out/c13_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c13_pytest.log
tail -6 gpurun_out/c13_pytest.log
