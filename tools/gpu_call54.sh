#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c54_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c54_pytest.log
tail -3 gpurun_out/c54_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg --no-roofline > gpurun_out/c54_bench.json 2> gpurun_out/c54_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c54_bench.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 300 python tools/timeline.py --tag c54 --pose off > gpurun_out/timeline_c54.log 2>&1; python tools/show_timeline.py gpurun_out/timeline_c54.csv 20 | grep "prop_"
