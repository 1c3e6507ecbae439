#!/bin/bash
# Runs the GPU test tiers one by one with their own timeouts so that a hanging
# kernel cannot take the whole call down; logs land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
run() { # name timeout cmd...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
TIERS=${TIERS:-"golden ops gemm ref model"}
for t in $TIERS; do
  case $t in
    golden) run golden 300 python -m pytest tests/test_gpu_golden.py -q -x -m gpu --durations=5;;
    ops)    run ops 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --durations=5;;
    gemm)   run gemm 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu --durations=5;;
    ref)    run ref 600 python -m pytest tests/test_gpu_ref_kernels.py -q -m gpu --durations=5;;
    model)  run model 900 python -m pytest tests/test_gpu_model.py -q -m gpu --durations=8;;
  esac
done
for f in $TIERS; do echo "--- $f"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/$f.log | head -30; done
