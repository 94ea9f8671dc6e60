#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 60 tools/atomics_probe > gpurun_out/c18_atomics_probe.log 2>&1
cat gpurun_out/c18_atomics_probe.log
SECONDS=0
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/c18_pytest.log 2>&1
echo "rc=$? wall=${SECONDS}s" >> gpurun_out/c18_pytest.log
tail -14 gpurun_out/c18_pytest.log
LINES_SHOWN=45 timeout 400 bash tools/launch_list.sh r2s8
