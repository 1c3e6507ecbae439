#!/bin/bash
# round 2, GPU call 6: TMA-store epilogue
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "bf16_shadow or full_size_logits or swish or affine or grouped" ) > gpurun_out/r6_gemm_a.log 2>&1
echo "rc=$?" >> gpurun_out/r6_gemm_a.log
( time timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_persistent.py tests/test_gpu_model.py -m gpu -x -q ) > gpurun_out/r6_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r6_tests.log
timeout 300 python scripts/gemm_stamps_bf16.py > gpurun_out/r6_stamps.txt 2>&1
( time MRN_GEMM_PROFILE_DUMP=gpurun_out/r6_gemm_bf16.csv timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r6_bench.log 2>&1
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "bf16" ) > gpurun_out/r6_fullsize.log 2>&1
echo "rc=$?" >> gpurun_out/r6_fullsize.log
tail -4 gpurun_out/r6_gemm_a.log; tail -6 gpurun_out/r6_tests.log; tail -3 gpurun_out/r6_bench.log | cut -c1-400; tail -5 gpurun_out/r6_fullsize.log; cut -c1-900 gpurun_out/r6_stamps.txt
