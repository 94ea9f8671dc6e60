#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
SECONDS=0
timeout 600 python bench.py > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
echo "bench rc=$? wall=${SECONDS}s"
python -c "
import json
d=json.load(open('gpurun_out/c20_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'], d['cpu_baseline'])
for o in d['roofline']['others']: print(o['kernel'][:60], round(o['launch_us'],1), round(o.get('frac'),4))
print(d['roofline']['frac'], [ (p['launch'],round(p['launch_us'],1),round(p['frac'],3)) for p in d['roofline']['per_launch']])"
tail -3 gpurun_out/c20_bench.err
SECONDS=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_field_fwd|k_field_bwd<|k_grid_bwd_run|k_grid_fwd_tmh_jac|k_prop_fwd|k_prop_bwd|k_step_losses|k_adam_flat' --launch-skip 78 --launch-count 13 -f -o gpurun_out/r02_step_kernels python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-schedule-leg > gpurun_out/c20_ncu.log 2>&1
echo "ncu rc=$? wall=${SECONDS}s"; ls -la gpurun_out/r02_step_kernels.ncu-rep
SECONDS=0
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-schedule-leg > gpurun_out/c20_bench_config3.json 2> gpurun_out/c20_bench_config3.err
echo "config3 rc=$? wall=${SECONDS}s"
python -c "
import json
d=json.load(open('gpurun_out/c20_bench_config3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for o in d['roofline']['others']: print(o['kernel'][:60], round(o['launch_us'],1), round(o.get('frac'),4))
print(d['roofline']['frac'], [ (p['launch'],round(p['launch_us'],1),round(p['frac'],3)) for p in d['roofline']['per_launch']])"
tail -3 gpurun_out/c20_bench_config3.err
