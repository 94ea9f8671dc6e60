#!/bin/bash
# where the backward's issue time goes: slot-copy request vs MMAs; wgrad / dgrad skipped; MMA rate probe with rotating wgrad slabs and smem noise
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for skip in 0 1 2 3; do
echo "== skip $skip"
timeout 120 tools/field_timing 4096 1 1 $skip > gpurun_out/c32_bwd_timing_$skip.log 2>&1; grep "backward rep" gpurun_out/c32_bwd_timing_$skip.log; sed -n '/tile 1 head L2/,/tile 1 base L0/p' gpurun_out/c32_bwd_timing_$skip.log | cut -c1-220
done
UMMA_RATE_FIRST=6 timeout 120 tools/umma_rate > gpurun_out/c32_umma_rate.log 2>&1; cat gpurun_out/c32_umma_rate.log
