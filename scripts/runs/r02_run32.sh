#!/bin/bash
# round 2, GPU call 32: bias-gradient sums outside the one-tile products (policy 1) / outside all products (policy 2)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "bias_gradient" ) > gpurun_out/r32_colsums.log 2>&1
echo "rc=$?" >> gpurun_out/r32_colsums.log; tail -3 gpurun_out/r32_colsums.log
( time MRN_COLSUM_POLICY=0 timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "bias_gradient" ) > gpurun_out/r32_colsums_p0.log 2>&1
echo "rc=$?" >> gpurun_out/r32_colsums_p0.log; tail -3 gpurun_out/r32_colsums_p0.log
( time timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_persistent.py tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q ) > gpurun_out/r32_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r32_tests.log; tail -4 gpurun_out/r32_tests.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic"
MRN_GEMM_PROFILE_DUMP=gpurun_out/r32_spans_p1.txt timeout 200 $B > gpurun_out/r32_p1.json 2> gpurun_out/r32_p1.err; echo "rc=$?"
MRN_COLSUM_POLICY=2 MRN_GEMM_PROFILE_DUMP=gpurun_out/r32_spans_p2.txt timeout 200 $B > gpurun_out/r32_p2.json 2> gpurun_out/r32_p2.err; echo "rc=$?"
MRN_SHADOW_TRACE=1 timeout 300 python bench.py --model s2s-deep-gru --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/r32_gru.json 2> gpurun_out/r32_gru.err; echo "rc=$?"
for f in p1 p2 gru; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r32_$f.json").read().strip().splitlines()[-1])
    print("$f", d["ms_per_step"], d.get("gpu_launches_per_step"), d.get("roofline",{}).get("gemm_ms_per_step"), d.get("roofline",{}).get("frac"), d.get("parity",{}).get("ok"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r32_$f.err").read()[-800:])
PY
done
for f in p1 p2; do python scripts/summarize_gemm_dump.py gpurun_out/r32_spans_$f.txt.spans | grep -E "NT|all launches" | head -10; done
grep "shadow-trace" gpurun_out/r32_gru.err | sort -t x -k2 -n | tail -8
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k "(transformer_base_full_size and (bf16 or replay)) or (deep_gru and (bf16 or replay))" ) > gpurun_out/r32_fullsize.log 2>&1
echo "rc=$?" >> gpurun_out/r32_fullsize.log; tail -5 gpurun_out/r32_fullsize.log
