#!/bin/bash
# round 2, GPU call 29: recurrent-path rework (lanes, GRU backward, one-launch concat, side-stream weight gradients) -
# operator / model / golden / reference-kernel tests, full-size deep-GRU parity (eager + replay), bench of config C with
# and without lanes, Transformer-base regression check, launch list of one eager deep-GRU step
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_golden.py tests/test_gpu_ref_kernels.py -m gpu -q -x ) > gpurun_out/r29_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r29_tests.log
( time timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k "deep_gru" ) > gpurun_out/r29_fullsize_gru.log 2>&1
echo "rc=$?" >> gpurun_out/r29_fullsize_gru.log
timeout 300 python bench.py --model s2s-deep-gru --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/r29_bench_gru.json 2> gpurun_out/r29_bench_gru.err
echo "bench gru rc=$?"
MRN_NO_LANES=1 timeout 300 python bench.py --model s2s-deep-gru --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-traffic > gpurun_out/r29_bench_gru_nolanes.json 2> gpurun_out/r29_bench_gru_nolanes.err
echo "bench gru nolanes rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic > gpurun_out/r29_bench_tb.json 2> gpurun_out/r29_bench_tb.err
echo "bench tb rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r29_launches_gru.csv python scripts/profile_step.py 2 4 s2s-deep-gru > gpurun_out/r29_profile_gru.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/r29_tests.log; tail -12 gpurun_out/r29_fullsize_gru.log
for f in gru gru_nolanes tb; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r29_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["ms_per_step"], d["value"], d.get("gpu_launches_per_step"), d.get("e2e",{}).get("ms_per_step"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r29_bench_$f.err").read()[-1500:])
PY
done
