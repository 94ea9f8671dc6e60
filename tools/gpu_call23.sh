#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_field_tc.py tests/test_exporter.py tests/test_full_size_parity.py tests/test_pose_opt.py -m gpu -q -x > gpurun_out/c23_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c23_pytest.log
tail -5 gpurun_out/c23_pytest.log
for f in 1 0; do
NVO_FIELD_BWD_PREFETCH=$f timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-schedule-leg > gpurun_out/c23_bench_p$f.json 2> gpurun_out/c23_bench_p$f.err
python -c "
import json
d=json.load(open('gpurun_out/c23_bench_p$f.json')); print('prefetch $f', d['value'], d['ms_per_step'], d['e2e']['value'])
for o in d['roofline']['others'][:4]: print('  ', o['kernel'][:60], round(o['launch_us'],1), round(o.get('frac'),4))"
tail -2 gpurun_out/c23_bench_p$f.err
done
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-schedule-leg > gpurun_out/c23_bench_config3.json 2> gpurun_out/c23_bench_config3.err
python -c "
import json
d=json.load(open('gpurun_out/c23_bench_config3.json')); print('config3', d['value'], d['ms_per_step'], d['e2e']['value'])
for o in d['roofline']['others'][:4]: print('  ', o['kernel'][:60], round(o['launch_us'],1), round(o.get('frac'),4))"
