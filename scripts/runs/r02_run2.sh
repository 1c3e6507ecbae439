#!/bin/bash
# round 2, GPU call 2: first run of the bf16-shadow GEMM mode (mode 4)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "bf16_shadow or full_size_logits" ) > gpurun_out/r2_gemm_a.log 2>&1
echo "rc=$?" >> gpurun_out/r2_gemm_a.log
( time timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q ) > gpurun_out/r2_gemm_b.log 2>&1
echo "rc=$?" >> gpurun_out/r2_gemm_b.log
( time timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x ) > gpurun_out/r2_model.log 2>&1
echo "rc=$?" >> gpurun_out/r2_model.log
( time MRN_GEMM_PROFILE_DUMP=gpurun_out/r2_gemm_bf16.csv timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --gemm-mode 4 ) > gpurun_out/r2_bench_bf16.log 2>&1
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "bf16 or tf32" ) > gpurun_out/r2_fullsize.log 2>&1
echo "rc=$?" >> gpurun_out/r2_fullsize.log
tail -5 gpurun_out/r2_gemm_a.log; tail -5 gpurun_out/r2_gemm_b.log; tail -5 gpurun_out/r2_model.log; tail -3 gpurun_out/r2_bench_bf16.log | cut -c1-1500; tail -5 gpurun_out/r2_fullsize.log
