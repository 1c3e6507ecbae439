#!/bin/bash
# round 2, GPU call 28: beam-search decoding on the GPU (n-best kernels, searches vs the oracle), then the whole gpu suite
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_translator.py -m gpu -x -q -s ) > gpurun_out/r28_translator.log 2>&1
echo "rc=$?" >> gpurun_out/r28_translator.log
( time timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_translator.py ) > gpurun_out/r28_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r28_tests.log
tail -40 gpurun_out/r28_translator.log; tail -6 gpurun_out/r28_tests.log
