#!/bin/bash
# round 2, GPU call 33: queued bias-gradient sums (issued behind the weight-gradient product) vs in-kernel sums; launch lists
# and ncu --set full captures of the final kernels (Transformer-base and config C, mode 4)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -k "bias_gradient" ) > gpurun_out/r33_colsums.log 2>&1
echo "rc=$?" >> gpurun_out/r33_colsums.log; tail -3 gpurun_out/r33_colsums.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic"
MRN_GEMM_PROFILE_DUMP=gpurun_out/r33_spans_p1.txt timeout 200 $B > gpurun_out/r33_p1.json 2> gpurun_out/r33_p1.err; echo "rc=$?"
MRN_COLSUM_POLICY=0 timeout 200 $B > gpurun_out/r33_p0.json 2> gpurun_out/r33_p0.err; echo "rc=$?"
for f in p1 p0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r33_$f.json").read().strip().splitlines()[-1])
    print("$f", d["ms_per_step"], d.get("gpu_launches_per_step"), d.get("roofline",{}).get("gemm_ms_per_step"), d.get("roofline",{}).get("frac"), d.get("parity",{}).get("ok"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r33_$f.err").read()[-800:])
PY
done
python scripts/summarize_gemm_dump.py gpurun_out/r33_spans_p1.txt.spans | grep -E "NT|all launches" | head -10
# launch lists (two eager steps each)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r33_launches_tb.csv python scripts/profile_step.py 2 4 > gpurun_out/r33_profile_tb.log 2>&1; echo "ncu tb rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r33_launches_gru.csv python scripts/profile_step.py 2 4 s2s-deep-gru > gpurun_out/r33_profile_gru.log 2>&1; echo "ncu gru rc=$?"
# full captures
for spec in "gGRUFastBackwardVec:200:2:s2s-deep-gru" "gColumnSumsBf16:30:3:transformer-base" "gCopyBlocks:20:2:s2s-deep-gru" "gAttVec:10:2:s2s-deep-gru"; do
  IFS=: read k skip cnt model <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -o gpurun_out/prof_$k -f python scripts/profile_step.py 2 4 $model > gpurun_out/r33_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
