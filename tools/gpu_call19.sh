#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python tools/overlap_probe.py > gpurun_out/c19_overlap.log 2>&1
cat gpurun_out/c19_overlap.log | tail -25
LINES_SHOWN=48 timeout 400 bash tools/launch_list.sh r2s8
tail -5 gpurun_out/ncu_bench_r2s8.log
