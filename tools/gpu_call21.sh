#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_full_size_parity.py -m gpu -q -x -k "loss or step or trajectory or config" > gpurun_out/c21_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c21_pytest.log
tail -5 gpurun_out/c21_pytest.log
for f in 1 0; do
NVO_LOSS_FAST=$f timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/c21_bench_f$f.json 2> gpurun_out/c21_bench_f$f.err
python -c "
import json
d=json.load(open('gpurun_out/c21_bench_f$f.json')); print('fast $f', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 300 python tools/timeline.py --tag r2s9 --pose off > gpurun_out/timeline_r2s9.log 2>&1
python tools/show_timeline.py gpurun_out/timeline_r2s9.csv 20
grep -n "k_step_losses" gpurun_out/timeline_r2s9.csv
