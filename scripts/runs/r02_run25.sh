#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "shared_a" ) > gpurun_out/r25_gemm.log 2>&1
echo "rc=$?" >> gpurun_out/r25_gemm.log
( time timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_persistent.py -m gpu -x -q ) > gpurun_out/r25_model.log 2>&1
echo "rc=$?" >> gpurun_out/r25_model.log
( time MRN_GEMM_PROFILE_DUMP=gpurun_out/r25_gemm_bf16.csv timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r25_bench.log 2>&1
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "bf16" ) > gpurun_out/r25_fullsize.log 2>&1
echo "rc=$?" >> gpurun_out/r25_fullsize.log
tail -5 gpurun_out/r25_gemm.log; tail -5 gpurun_out/r25_model.log; tail -3 gpurun_out/r25_bench.log | cut -c1-300; tail -4 gpurun_out/r25_fullsize.log
