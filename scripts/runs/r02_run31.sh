#!/bin/bash
# round 2, GPU call 31: why are the K-major (NT) input-gradient products 2.5x slower per k-block than the NN ones?
# (a) without the fused bias-gradient sums, (b) 128-wide tiles, (c) split-K 2 for the 3200x512x2048 product
mkdir -p gpurun_out
B="python bench.py --steps 15 --warmup 5 --no-cpu-baseline --no-traffic --no-parity"
MRN_NO_FUSE_BIAS_GRAD=1 MRN_GEMM_PROFILE_DUMP=gpurun_out/r31_spans_nofuse.txt timeout 200 $B > gpurun_out/r31_nofuse.json 2> gpurun_out/r31_nofuse.err; echo "rc=$?"
MRN_GEMM_TRY=3200,512,2048,1.0,128,1 MRN_GEMM_PROFILE_DUMP=gpurun_out/r31_spans_bn128.txt timeout 200 $B > gpurun_out/r31_bn128.json 2> gpurun_out/r31_bn128.err; echo "rc=$?"
MRN_GEMM_TRY=3200,512,2048,1.0,64,2 MRN_GEMM_PROFILE_DUMP=gpurun_out/r31_spans_sp2.txt timeout 200 $B > gpurun_out/r31_sp2.json 2> gpurun_out/r31_sp2.err; echo "rc=$?"
for f in nofuse bn128 sp2; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r31_$f.json").read().strip().splitlines()[-1])
    print("$f", d["ms_per_step"], d.get("gpu_launches_per_step"), d.get("roofline",{}).get("gemm_ms_per_step"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r31_$f.err").read()[-800:])
PY
python scripts/summarize_gemm_dump.py gpurun_out/r31_spans_$f.txt.spans | grep -E "NT|all launches" | head -12
done
