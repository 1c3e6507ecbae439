#!/bin/bash
# round 2, GPU call 27: verification of HEAD (all gpu tests, default bench with the in-run traffic child, reference arm, smoke)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r27_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r27_tests.log
( time timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r27_bench.log 2>&1
echo "rc=$?" >> gpurun_out/r27_bench.log
( time timeout 400 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/r27_ref.log 2>&1
echo "rc=$?" >> gpurun_out/r27_ref.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r27_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/r27_smoke.log
tail -6 gpurun_out/r27_tests.log; tail -4 gpurun_out/r27_bench.log | cut -c1-1500; tail -3 gpurun_out/r27_ref.log | cut -c1-600; tail -3 gpurun_out/r27_smoke.log
