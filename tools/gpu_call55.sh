#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c55_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c55_pytest.log
tail -3 gpurun_out/c55_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg > gpurun_out/c55_bench.json 2> gpurun_out/c55_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c55_bench.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], [ (round(x['launch_us'],1)) for x in d['roofline'].get('per_launch', [])])
for o in d['roofline']['others'][:4]: print('  ', o['kernel'][:60], round(o['launch_us'],1), round(o.get('frac'),4))"
