#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s ) > gpurun_out/m2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/m2_tests.log
for sc in weak strong; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --scaling $sc ) > gpurun_out/m2_bench_$sc.log 2>&1
done
tail -15 gpurun_out/m2_tests.log | cut -c1-1200; for sc in weak strong; do grep '^{' gpurun_out/m2_bench_$sc.log | cut -c1-400; done
