#!/bin/bash
# Runs the GPU test tiers one by one with their own timeouts so that a hanging
# kernel cannot take the whole call down; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name timeout cmd...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run golden 300 python -m pytest tests/test_gpu_golden.py -q -x -m gpu --durations=8
run ops 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --durations=8
run gemm 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --durations=8
run model 900 python -m pytest tests/test_gpu_model.py -q -m gpu --durations=8
for f in golden ops gemm model; do echo "--- $f"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/$f.log | head -40; done
