#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; ( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 "$@" ) > gpurun_out/m8_$name.log 2>&1; grep '^{' gpurun_out/m8_$name.log | cut -c1-260; }
run weak
MRN_NO_EXCHANGE_OVERLAP=1 run weak_nooverlap
run strong --scaling strong
run big_strong --model transformer-big --scaling strong
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k "8" ) > gpurun_out/m8_tests.log 2>&1
tail -4 gpurun_out/m8_tests.log | cut -c1-300
tail -5 gpurun_out/m8_big_strong.log | grep -v '^{' | cut -c1-300
