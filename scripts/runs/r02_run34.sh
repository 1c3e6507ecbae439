#!/bin/bash
# round 2, GPU call 34 (2 GPUs): the multi-GPU paths after the device-layer changes (per-lane side streams, lanes, queued
# bias-gradient sums): exchanges vs the oracle's SyncGraphGroup, AsyncGraphGroup on two GPUs, weak / strong bench lines,
# config C with the split step (lanes closed and re-opened around the exchange hook)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s ) > gpurun_out/r34_multi_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r34_multi_tests.log; grep -E "passed|failed|skipped|rc=" gpurun_out/r34_multi_tests.log | tail -3
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-traffic --no-parity > gpurun_out/r34_weak2.json 2> gpurun_out/r34_weak2.err; echo "weak rc=$?"
timeout 300 $T --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --scaling strong --no-cpu-baseline --no-traffic --no-parity > gpurun_out/r34_strong2.json 2> gpurun_out/r34_strong2.err; echo "strong rc=$?"
timeout 300 $T --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 --model s2s-deep-gru --no-cpu-baseline --no-traffic --no-parity > gpurun_out/r34_gru2.json 2> gpurun_out/r34_gru2.err; echo "gru rc=$?"
for f in weak2 strong2 gru2; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r34_$f.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$f", d["n_gpus"], d["ms_per_step"], d["value"], d.get("scaling"), d.get("config",{}).get("exchange"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r34_$f.err").read()[-1500:])
PY
done
