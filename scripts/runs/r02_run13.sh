#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_persistent.py -m gpu -x -q ) > gpurun_out/r13_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r13_tests.log
timeout 300 python scripts/gemm_stamps_bf16.py 2>&1 | head -1 | cut -c1-900 > gpurun_out/r13_stamps.txt
( time MRN_GEMM_PROFILE_DUMP=gpurun_out/r13_gemm_bf16.csv timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r13_bench.log 2>&1
( time timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q ) > gpurun_out/r13_model.log 2>&1
echo "rc=$?" >> gpurun_out/r13_model.log
tail -5 gpurun_out/r13_tests.log; cat gpurun_out/r13_stamps.txt; tail -3 gpurun_out/r13_bench.log | cut -c1-300; tail -5 gpurun_out/r13_model.log
