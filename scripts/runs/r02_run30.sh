#!/bin/bash
# round 2, GPU call 30: bias-gradient column sums dealt out over tile columns (both bf16 kernels), eight-stage ring for
# one-tile-row products, per-lane side streams, bf16 copies from concat / kept-axis sums - tests first (hang guard), then benches
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k "bias_gradient" ) > gpurun_out/r30_colsums.log 2>&1
echo "rc=$?" >> gpurun_out/r30_colsums.log
tail -4 gpurun_out/r30_colsums.log
( time timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_persistent.py tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x ) > gpurun_out/r30_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r30_tests.log
tail -4 gpurun_out/r30_tests.log
MRN_GEMM_PROFILE_DUMP=gpurun_out/r30_gemm_spans_tb.txt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic > gpurun_out/r30_bench_tb.json 2> gpurun_out/r30_bench_tb.err
echo "bench tb rc=$?"
MRN_SHADOW_TRACE=1 MRN_GEMM_PROFILE_DUMP=gpurun_out/r30_gemm_spans_gru.txt timeout 300 python bench.py --model s2s-deep-gru --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/r30_bench_gru.json 2> gpurun_out/r30_bench_gru.err
echo "bench gru rc=$?"
MRN_GEMM_NO_DEEP_RING=1 timeout 300 python bench.py --model s2s-deep-gru --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/r30_bench_gru_ring4.json 2> gpurun_out/r30_bench_gru_ring4.err
echo "bench gru ring4 rc=$?"
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k "(transformer_base_full_size and (bf16 or replay)) or deep_gru_full_size_replay or (deep_gru and bf16)" ) > gpurun_out/r30_fullsize.log 2>&1
echo "rc=$?" >> gpurun_out/r30_fullsize.log
tail -8 gpurun_out/r30_fullsize.log
for f in tb gru gru_ring4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r30_bench_$f.json").read().strip().splitlines()[-1])
    r=d.get("roofline",{})
    print("$f", d["ms_per_step"], d["value"], d.get("gpu_launches_per_step"), "gemm ms", r.get("gemm_ms_per_step"), "frac", r.get("frac"), d.get("parity",{}).get("ok"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r30_bench_$f.err").read()[-1500:])
PY
done
grep "shadow-trace" gpurun_out/r30_bench_gru.err | sort -t x -k2 -n | tail -20
