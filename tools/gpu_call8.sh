#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_trainer.py tests/test_full_size_parity.py -m gpu -q -x > gpurun_out/c8_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c8_pytest.log
tail -8 gpurun_out/c8_pytest.log
timeout 300 python tools/timeline.py --tag r2s3 > gpurun_out/timeline_r2s3.log 2>&1
tail -3 gpurun_out/timeline_r2s3.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
python -c "
import json; d=json.load(open('gpurun_out/c8_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"
