#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for ns in 0 200 1000 5000; do
echo "== NVO_FIELD_BWD_WAIT_NS=$ns"
NVO_FIELD_BWD_WAIT_NS=$ns timeout 120 tools/field_timing 4096 1 1 > gpurun_out/c30_bwd_timing_$ns.log 2>&1; grep "backward rep" gpurun_out/c30_bwd_timing_$ns.log; sed -n '/tile 1 head L2/,/tile 1 base L0/p' gpurun_out/c30_bwd_timing_$ns.log | cut -c1-200
done
