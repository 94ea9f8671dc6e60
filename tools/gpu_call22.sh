#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_bwd|k_step_losses|k_weights_bwd|k_render_fwd|k_pdf_resample|k_weights_fwd' --launch-skip 96 --launch-count 14 -f -o gpurun_out/r02_step_kernels_b python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/c22_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_step_kernels_b.ncu-rep
