#!/bin/bash
# round 2, GPU call 35: verification of the final code - whole gpu suite, smoke, the bench lines (default, reference arm,
# config C, padded batches) and the stream-priority A/B
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r35_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r35_tests.log; grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/r35_tests.log | tail -8
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r35_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r35_smoke.log | cut -c1-200
MRN_GEMM_PROFILE_DUMP=gpurun_out/r35_spans.txt timeout 600 python bench.py > gpurun_out/r35_bench.json 2> gpurun_out/r35_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r35_reference.json 2> gpurun_out/r35_reference.err; echo "reference rc=$?"
MRN_NO_STREAM_PRIORITY=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --no-parity > gpurun_out/r35_noprio.json 2> gpurun_out/r35_noprio.err; echo "noprio rc=$?"
timeout 400 python bench.py --model s2s-deep-gru --steps 10 --warmup 3 --no-cpu-baseline --no-traffic > gpurun_out/r35_gru.json 2> gpurun_out/r35_gru.err; echo "gru rc=$?"
timeout 200 python bench.py --padded --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --no-parity > gpurun_out/r35_padded.json 2> gpurun_out/r35_padded.err; echo "padded rc=$?"
for f in bench reference noprio gru padded; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r35_$f.json").read().strip().splitlines() if l.startswith("{")][-1])
    r=d.get("roofline",{})
    print("$f", d.get("ms_per_step"), d.get("value"), d.get("gpu_launches_per_step"), "gemm", r.get("gemm_ms_per_step"), r.get("frac"), "traffic", r.get("traffic"), "parity", d.get("parity",{}).get("ok"), d.get("parity",{}).get("logits_max_rel_err"), "e2e", d.get("e2e",{}).get("value"), "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r35_$f.err").read()[-1000:])
PY
done
python scripts/summarize_gemm_dump.py gpurun_out/r35_spans.txt.spans | grep -E "NT|all launches" | head -10
