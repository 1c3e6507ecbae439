#!/bin/bash
# round 2, GPU call 38: sanity of the final binary (host-side changes since run 37: queued sums dropped after an unfinished
# sweep, stream priorities opt-in): smoke, a slice of the suite, one short bench
mkdir -p gpurun_out
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r38_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_translator.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-traffic > gpurun_out/r38_bench.json 2> gpurun_out/r38_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r38_bench.json
