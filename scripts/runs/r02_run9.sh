#!/bin/bash
# round 2, GPU call 9: shadow-only tensors
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_persistent.py -m gpu -x -q ) > gpurun_out/r9_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r9_tests.log
( time MRN_GEMM_PROFILE_DUMP=gpurun_out/r9_gemm_bf16.csv timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r9_bench.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -q -x ) > gpurun_out/r9_tests2.log 2>&1
echo "rc=$?" >> gpurun_out/r9_tests2.log
tail -12 gpurun_out/r9_tests.log; tail -3 gpurun_out/r9_bench.log | cut -c1-400; tail -5 gpurun_out/r9_tests2.log
