#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer.py tests/test_checkpoint.py tests/test_exchange.py -m gpu -q -x > gpurun_out/c24_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c24_pytest.log
tail -6 gpurun_out/c24_pytest.log
for f in off on; do
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-roofline --no-schedule-leg --defer-fields $f > gpurun_out/c24_bench_$f.json 2> gpurun_out/c24_bench_$f.err
python -c "
import json
d=json.load(open('gpurun_out/c24_bench_$f.json')); print('defer $f', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'], d['final_loss'])"
tail -2 gpurun_out/c24_bench_$f.err
done
