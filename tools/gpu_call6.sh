#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_field_tc.py -m gpu -q > gpurun_out/c6_field_tc.log 2>&1
echo "rc=$?" >> gpurun_out/c6_field_tc.log
tail -25 gpurun_out/c6_field_tc.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c6_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c6_pytest.log
tail -25 gpurun_out/c6_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
python -c "
import json; d=json.load(open('gpurun_out/c6_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"
