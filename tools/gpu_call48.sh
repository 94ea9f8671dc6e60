#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_exchange.py -m gpu -q -x -k deferred > gpurun_out/c48_pytest.log 2>&1
grep -v "^$" gpurun_out/c48_pytest.log | grep -i "error\|assert\|Traceback\|raise\|File\|E  " | head -40
tail -5 gpurun_out/c48_pytest.log
