#!/bin/bash
# round 2, call 1: full GPU parity suite (incl. the new full-size oracle tests) + one bench line per BASELINE config at N=1
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
nproc >> gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_config2.json 2> gpurun_out/c1_bench_config2.err
timeout 300 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-schedule-leg > gpurun_out/c1_bench_config3.json 2> gpurun_out/c1_bench_config3.err
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-schedule-leg > gpurun_out/c1_bench_config4_n1.json 2> gpurun_out/c1_bench_config4_n1.err
timeout 300 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/c1_bench_config5_n1.json 2> gpurun_out/c1_bench_config5_n1.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c1_bench_reference.json 2> gpurun_out/c1_bench_reference.err
tail -3 gpurun_out/c1_pytest.log
