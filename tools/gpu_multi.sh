#!/bin/bash
# usage: tools/gpu_multi.sh N tag [configs...]  -- bench lines of configs 2, 4, 5 on N GPUs of one box (torchrun)
N=$1; tag=$2; shift 2
CONFIGS=${@:-2 4 5}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${tag}_smi.txt
port=29511
for c in $CONFIGS; do
  port=$((port+1))
  case $c in
    2) extra="--steps 40 --warmup 5 --no-cpu-baseline --no-schedule-leg";;
    4) extra="--config 4 --steps 8 --warmup 3 --no-cpu-baseline --no-schedule-leg";;
    5) extra="--config 5 --steps 5 --warmup 3";;
    3) extra="--config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-schedule-leg";;
  esac
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N $extra > gpurun_out/${tag}_config$c.json 2> gpurun_out/${tag}_config$c.err
  echo "config $c rc=$? bytes=$(wc -c < gpurun_out/${tag}_config$c.json)"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_config$c.json') if l.startswith('{')][-1]); print('config$c', d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], d.get('exchange'))"
  tail -n 3 gpurun_out/${tag}_config$c.err
done
