#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for skip in 0 3; do
echo "== skip $skip"
timeout 120 tools/field_timing 4096 1 1 $skip > gpurun_out/c33_bwd_timing_$skip.log 2>&1; grep "backward rep" gpurun_out/c33_bwd_timing_$skip.log; sed -n '/tile 1 head L2/,/tile 2 base L0/p' gpurun_out/c33_bwd_timing_$skip.log | cut -c1-220
done
