#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 tools/field_timing 65536 1 1 > gpurun_out/c51_bwd_timing_65k.log 2>&1; grep "rep" gpurun_out/c51_bwd_timing_65k.log; sed -n '/per CTA, ns/,/mean setup/p' gpurun_out/c51_bwd_timing_65k.log | tail -3; sed -n '/tile 4 head L2/,/tile 7 base L0/p' gpurun_out/c51_bwd_timing_65k.log | cut -c1-250
NVO_FIELD_BWD_PREFETCH=0 timeout 120 tools/field_timing 65536 1 1 > gpurun_out/c51_bwd_timing_65k_nopf.log 2>&1; grep "backward rep" gpurun_out/c51_bwd_timing_65k_nopf.log
