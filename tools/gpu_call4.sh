#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_field_tc.py -m gpu -q -x > gpurun_out/c4_field_tc.log 2>&1
echo "rc=$?" >> gpurun_out/c4_field_tc.log
tail -5 gpurun_out/c4_field_tc.log
timeout 300 python tools/field_tc_bench.py 4096 65536 > gpurun_out/c4_field_bench.log 2>&1
tail -n 3 gpurun_out/c4_field_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_field_fwd -s 3 -c 2 -o gpurun_out/prof_field_fwd_4096 python tools/field_tc_bench.py 4096 > gpurun_out/c4_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_field_fwd -s 3 -c 1 -o gpurun_out/prof_field_fwd_65536 python tools/field_tc_bench.py 65536 > gpurun_out/c4_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
