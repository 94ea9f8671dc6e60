#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_trainer.py -m gpu -q -x > gpurun_out/c9_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c9_pytest.log
tail -4 gpurun_out/c9_pytest.log
timeout 300 python tools/timeline.py --tag r2s4 > gpurun_out/timeline_r2s4.log 2>&1
tail -2 gpurun_out/timeline_r2s4.log
NVO_EARLY_FIELDS_OPT=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/c9_bench_noearly.json 2> gpurun_out/c9_bench_noearly.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
python -c "
import json
for f in ('c9_bench_noearly','c9_bench'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d.get('reference_schedule',{}).get('value'), d['e2e']['value'])"
