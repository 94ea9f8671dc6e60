#!/bin/bash
# usage: tools/gpu_multi2.sh N tag  -- config 2 with / without the deferred fields update, then the exchange test at N ranks
N=$1; tag=$2
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
port=29611
for f in on off; do
  port=$((port+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-schedule-leg --defer-fields $f > gpurun_out/${tag}_defer_$f.json 2> gpurun_out/${tag}_defer_$f.err
  echo "defer $f rc=$?"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}_defer_$f.json') if l.startswith('{')][-1]); print('defer $f', d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['final_loss'], d.get('exchange'))"
  tail -n 2 gpurun_out/${tag}_defer_$f.err
done
