#!/bin/bash
# usage: tools/gpu_multi4.sh N tag [tests]  -- the driver's command line at N ranks (config 2, defaults), optionally the 2-rank GPU tests first
N=$1; tag=$2
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
if [ "$3" = "tests" ]; then
  timeout 600 python -m pytest tests/test_exchange.py tests/test_sharding_gloo.py -m gpu -q -x 2>&1 | tail -3
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-schedule-leg > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/${tag}.json') if l.startswith('{')][-1]); print('N', d['n_gpus'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['final_loss'], d.get('exchange', {}).get('us_per_step'), d.get('replicas_bit_identical'), d.get('config'))"
tail -n 2 gpurun_out/${tag}.err
