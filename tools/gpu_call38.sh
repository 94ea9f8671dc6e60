#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 tools/field_timing 4096 1 1 > gpurun_out/c38_bwd_timing.log 2>&1; grep "backward rep" gpurun_out/c38_bwd_timing.log; sed -n '/tile 0 head L2/,/tile 3 base L0/p' gpurun_out/c38_bwd_timing.log | cut -c1-250
