#!/bin/bash
mkdir -p gpurun_out
for blocks in 296 74 1184; do
  echo "== background blocks $blocks"
  MRN_EXCHANGE_BLOCKS=$blocks timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | grep '^{' | cut -c1-250
done
echo "== strong"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --scaling strong 2>&1 | grep '^{' | cut -c1-250
