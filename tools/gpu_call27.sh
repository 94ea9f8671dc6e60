#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c27_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c27_pytest.log
tail -4 gpurun_out/c27_pytest.log
for f in 1 0; do
NVO_RAYS_LANE_BLOCKED=$f timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/c27_bench_lb$f.json 2> gpurun_out/c27_bench_lb$f.err
python -c "
import json
d=json.load(open('gpurun_out/c27_bench_lb$f.json')); print('lane-blocked $f', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 300 python tools/timeline.py --tag r2s10 --pose off > gpurun_out/timeline_r2s10.log 2>&1
python tools/show_timeline.py gpurun_out/timeline_r2s10.csv 3 | grep -E "k_weights|k_pdf|k_render|k_step|k_sample"
