#!/bin/bash
# static issue schedule of k_field_bwd: phase timing (static vs dynamic), field tests, full GPU suite, bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for st in 1 0; do
echo "== NVO_FIELD_BWD_STATIC=$st"
NVO_FIELD_BWD_STATIC=$st timeout 120 tools/field_timing 4096 1 1 > gpurun_out/c34_bwd_timing_$st.log 2>&1; grep "backward rep" gpurun_out/c34_bwd_timing_$st.log; sed -n '/tile 1 head L2/,/tile 1 base L0/p' gpurun_out/c34_bwd_timing_$st.log | cut -c1-220
done
timeout 600 python -m pytest tests/test_field_tc.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c34_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c34_pytest.log
tail -4 gpurun_out/c34_pytest.log
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-schedule-leg > gpurun_out/c34_bench.json 2> gpurun_out/c34_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c34_bench.json')); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])
for o in d['roofline']['others'][:4]: print('  ', o['kernel'][:60], round(o['launch_us'],1), round(o.get('frac'),4))"
