#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_field_tc.py -m gpu -q -x > gpurun_out/c3_field_tc.log 2>&1
echo "rc=$?" >> gpurun_out/c3_field_tc.log
tail -25 gpurun_out/c3_field_tc.log
NVO_FIELD_GROUPS=3 timeout 300 python tools/field_tc_bench.py 4096 65536 > gpurun_out/c3_field_bench_g3.log 2>&1
NVO_FIELD_GROUPS=2 timeout 300 python tools/field_tc_bench.py 4096 65536 > gpurun_out/c3_field_bench_g2.log 2>&1
tail -3 gpurun_out/c3_field_bench_g3.log gpurun_out/c3_field_bench_g2.log
timeout 600 python -m pytest tests/test_full_size_parity.py tests/test_trainer.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/c3_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c3_pytest.log
tail -15 gpurun_out/c3_pytest.log
