#!/bin/bash
cd "$GRAFT_REPO_ROOT/_old"
mkdir -p ../gpurun_out
timeout 600 python -m pytest tests/test_exchange.py -m gpu -q -x -k deferred > ../gpurun_out/c49_pytest_old.log 2>&1
grep "Index\|^E     [0-9]\|passed\|failed" ../gpurun_out/c49_pytest_old.log | head -12
