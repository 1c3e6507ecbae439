#!/bin/bash
# round 2, GPU call 36: the two decoder tests that failed in run 35 (+ the whole translator file and the lane test), repeated three times
mkdir -p gpurun_out
for i in 1 2 3; do
( time timeout 600 python -m pytest tests/test_gpu_translator.py -m gpu -q ) > gpurun_out/r36_translator_$i.log 2>&1
echo "rc=$?" >> gpurun_out/r36_translator_$i.log; grep -E "passed|failed|rc=|^FAILED" gpurun_out/r36_translator_$i.log | tail -4
done
( time timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -k "lanes or replay" ) > gpurun_out/r36_model.log 2>&1
echo "rc=$?" >> gpurun_out/r36_model.log; grep -E "passed|failed|rc=|^FAILED" gpurun_out/r36_model.log | tail -4
