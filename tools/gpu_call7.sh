#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c7_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c7_pytest.log
tail -8 gpurun_out/c7_pytest.log
timeout 300 python tools/timeline.py --tag r2s2 > gpurun_out/timeline_r2s2.log 2>&1
tail -30 gpurun_out/timeline_r2s2.log
timeout 600 bash tools/launch_list.sh r2s2
