#!/bin/bash
# usage: tools/gpu_multi3.sh N tag  -- config 2, deferred fields update with 1 / 2 / all CTAs per SM for the exchange
N=$1; tag=$2
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
port=29711
for c in 1 2 0; do
  port=$((port+1))
  NVO_DEFER_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline --no-schedule-leg --defer-fields on > gpurun_out/${tag}_ctas$c.json 2> gpurun_out/${tag}_ctas$c.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_ctas$c.json') if l.startswith('{')][-1]); print('ctas $c', d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['final_loss'], d['exchange']['us_per_step'])"
  tail -n 1 gpurun_out/${tag}_ctas$c.err
done
