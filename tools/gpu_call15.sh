#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_full_size_parity.py tests/test_trainer.py tests/test_exchange.py tests/test_pose_opt.py -m gpu -q -x > gpurun_out/c15_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c15_pytest.log
tail -4 gpurun_out/c15_pytest.log
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
python -c "
import json
d=json.load(open('gpurun_out/c15_bench.json')); print(d['value'], d['ms_per_step'], d.get('reference_schedule',{}).get('value'), d['e2e']['value'], d['gpu_launches_per_step'], d['pose_opt'])
for o in d['roofline']['others']: print(o['kernel'][:50], o['launch_us'], o.get('frac'))
print(d['roofline']['frac'], [ (p['launch'],p['launch_us'],p['frac']) for p in d['roofline']['per_launch']])"
tail -3 gpurun_out/c15_bench.err
