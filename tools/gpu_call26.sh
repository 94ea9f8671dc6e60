#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-schedule-leg > gpurun_out/c26_$tag.json 2> gpurun_out/c26_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/c26_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],4), [ (round(p['launch_us'],1)) for p in d['roofline']['per_launch']], round(d['roofline']['frac'],3))"
  tail -1 gpurun_out/c26_$tag.err
}
run base NVO_GRID_BWD_ROLLED=0
run rolled1_g16 NVO_GRID_BWD_ROLLED=1 NVO_GRID_BWD_ROLLED_G=16
run rolled1_g8 NVO_GRID_BWD_ROLLED=1 NVO_GRID_BWD_ROLLED_G=8
run rolled1_g32 NVO_GRID_BWD_ROLLED=1 NVO_GRID_BWD_ROLLED_G=32
run rolled2_g16 NVO_GRID_BWD_ROLLED=2 NVO_GRID_BWD_ROLLED_G=16
NVO_GRID_BWD_ROLLED=2 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -q -x -k "grid or scatter or table" 2>&1 | tail -3
