#!/bin/bash
# ncu launch list of the mapping step (eager, serialised, cold cache): per-kernel alone-times. usage: tools/launch_list.sh <tag> [extra bench args]
tag=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 420 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-schedule-leg "$@" > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_$tag.csv > gpurun_out/launches_$tag.md
head -${LINES_SHOWN:-24} gpurun_out/launches_$tag.md
