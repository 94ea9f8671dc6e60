#!/bin/bash
# 2-GPU validation of the fused peer-memory exchange (run under gpurun --gpus 2)
set -x
nvidia-smi topo -m 2>&1 | head -8
timeout 300 python -m pytest tests/test_exchange.py -m gpu -x -q 2>&1 | tail -5
for arm in fused nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 --exchange $arm 2>&1 | tail -2
done
