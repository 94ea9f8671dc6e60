#!/bin/bash
# final evidence of the build in the tree: tests, default bench, config 3, launch list, full ncu capture of the step's kernels, timeline
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1
echo "rc=$? wall=${SECONDS}s" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python -c "
import json
d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'], d['cpu_baseline']['value'], d['reference_schedule']['value'], d['pose_opt']['value'], d['roofline']['frac'])
for o in d['roofline']['others'][:4]: print('  ', o['kernel'][:60], round(o['launch_us'],1), round(o.get('frac'),4))"
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-schedule-leg > gpurun_out/final_bench_config3.json 2> gpurun_out/final_bench_config3.err
python -c "
import json
d=json.load(open('gpurun_out/final_bench_config3.json')); print('config3', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
for o in d['roofline']['others'][:4]: print('  ', o['kernel'][:60], round(o['launch_us'],1), round(o.get('frac'),4))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 400 gpurun_out/final_bench_reference.json
LINES_SHOWN=4 timeout 400 bash tools/launch_list.sh final
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_fwd|k_field_bwd|k_grid_bwd_run|k_grid_fwd_tmh_jac|k_prop_fwd|k_prop_bwd|k_step_losses|k_adam_flat|k_weights|k_pdf_resample' --launch-skip 140 --launch-count 22 -f -o gpurun_out/r02_final_step_kernels python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/final_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_final_step_kernels.ncu-rep
timeout 300 python tools/timeline.py --tag final --pose off > gpurun_out/timeline_final.log 2>&1; tail -1 gpurun_out/timeline_final.log
