#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_field_tc.py -m gpu -q -x > gpurun_out/c5_field_tc.log 2>&1
echo "rc=$?" >> gpurun_out/c5_field_tc.log
tail -15 gpurun_out/c5_field_tc.log
timeout 120 build/tools/field_timing 4096 1 > gpurun_out/c5_timing_4096_save.log 2>&1
timeout 120 build/tools/field_timing 4096 0 > gpurun_out/c5_timing_4096_nosave.log 2>&1
timeout 120 build/tools/field_timing 65536 1 > gpurun_out/c5_timing_65536_save.log 2>&1
head -50 gpurun_out/c5_timing_4096_save.log
timeout 120 build/tools/umma_rate > gpurun_out/c5_umma_rate.log 2>&1
cat gpurun_out/c5_umma_rate.log
timeout 600 python -m pytest tests/test_full_size_parity.py tests/test_trainer.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/c5_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c5_pytest.log
tail -15 gpurun_out/c5_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
cat gpurun_out/c5_bench.json
