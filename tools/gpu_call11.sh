#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_field_tc.py -m gpu -q -x > gpurun_out/c11_field_tc.log 2>&1
echo "rc=$?" >> gpurun_out/c11_field_tc.log
tail -3 gpurun_out/c11_field_tc.log
timeout 20 build/tools/field_timing 4096 1 > gpurun_out/c11_timing.log 2>&1; head -14 gpurun_out/c11_timing.log
NVO_FIELD_GROUPS=1 timeout 20 build/tools/field_timing 4096 1 > gpurun_out/c11_timing_g1.log 2>&1; head -14 gpurun_out/c11_timing_g1.log
timeout 100 python tools/field_tc_bench.py 4096 65536 > gpurun_out/c11_field_bench.log 2>&1
python - <<PY
import json
for r in json.load(open("gpurun_out/field_tc_bench.json")):
    print(r["rays"], "save: %.1f us %.3f | nosave: %.1f us %.3f" % (r["fused_fwd_us_saving"], r["fused_fwd_frac_of_tensor_peak_saving"], r["fused_fwd_us"], r["fused_fwd_frac_of_tensor_peak"]))
PY
