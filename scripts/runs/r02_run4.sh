#!/bin/bash
# round 2, GPU call 4: persistent GEMM
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q -k "bf16_shadow or full_size_logits or swish" ) > gpurun_out/r4_gemm_a.log 2>&1
echo "rc=$?" >> gpurun_out/r4_gemm_a.log
( time timeout 900 python -m pytest tests/test_gpu_persistent.py -m gpu -x -q ) > gpurun_out/r4_persist.log 2>&1
echo "rc=$?" >> gpurun_out/r4_persist.log
( time timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x ) > gpurun_out/r4_model.log 2>&1
echo "rc=$?" >> gpurun_out/r4_model.log
( time MRN_GEMM_PROFILE_DUMP=gpurun_out/r4_gemm_bf16.csv timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r4_bench.log 2>&1
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "bf16" ) > gpurun_out/r4_fullsize.log 2>&1
echo "rc=$?" >> gpurun_out/r4_fullsize.log
tail -4 gpurun_out/r4_gemm_a.log; tail -12 gpurun_out/r4_persist.log; tail -5 gpurun_out/r4_model.log; tail -3 gpurun_out/r4_bench.log | cut -c1-600; tail -5 gpurun_out/r4_fullsize.log
