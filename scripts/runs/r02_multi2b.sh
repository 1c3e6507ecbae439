#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s ) > gpurun_out/m2b_tests.log 2>&1
echo "rc=$?" >> gpurun_out/m2b_tests.log
for sc in weak strong; do
  ( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --scaling $sc ) > gpurun_out/m2b_bench_$sc.log 2>&1
done
( time MRN_NO_EXCHANGE_OVERLAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 ) > gpurun_out/m2b_bench_weak_nooverlap.log 2>&1
tail -12 gpurun_out/m2b_tests.log | cut -c1-600; for sc in weak strong weak_nooverlap; do grep '^{' gpurun_out/m2b_bench_$sc.log | cut -c1-330; tail -3 gpurun_out/m2b_bench_$sc.log | grep -v "^{" | cut -c1-300; done
