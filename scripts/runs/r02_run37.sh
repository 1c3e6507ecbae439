#!/bin/bash
# round 2, GPU call 37: whole gpu suite on the final code, the copy-task decoder test five more times, smoke, default bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r37_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r37_tests.log; grep -E "passed|failed|rc=|^FAILED|^ERROR" gpurun_out/r37_tests.log | tail -6
for i in 1 2 3 4 5; do timeout 200 python -m pytest tests/test_gpu_translator.py -m gpu -q -k "copy_task or nbest_lists" 2>&1 | tail -1; done
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r37_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/r37_bench.json 2> gpurun_out/r37_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r37_bench.json
