#!/bin/bash
mkdir -p gpurun_out
( time MRN_GEMM_PROFILE_DUMP=gpurun_out/r14_gemm_bf16.csv timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r14_bench.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r14_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r14_pytest.log
( time timeout 600 python bench.py --model s2s-deep-gru --steps 10 --warmup 3 ) > gpurun_out/r14_bench_gru.log 2>&1
( time timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --padded --no-cpu-baseline ) > gpurun_out/r14_bench_padded.log 2>&1
tail -3 gpurun_out/r14_bench.log | cut -c1-300; tail -6 gpurun_out/r14_pytest.log; tail -3 gpurun_out/r14_bench_gru.log | cut -c1-300; tail -3 gpurun_out/r14_bench_padded.log | cut -c1-300
