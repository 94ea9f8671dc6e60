#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python tools/diag_grad_parity.py 4096 fp32 > gpurun_out/c2_diag.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
tail -30 gpurun_out/c2_pytest.log
