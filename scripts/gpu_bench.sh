#!/bin/bash
# bench line + op table + ncu launch list + one full capture of the tensor-core GEMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "$OPBENCH" ]; then timeout 900 python scripts/op_bench.py > gpurun_out/op_table.json 2> gpurun_out/op_table.err; echo "op_bench rc=$?"; tail -3 gpurun_out/op_table.err; fi
if [ -n "$NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 2 1 > gpurun_out/profile_step.log 2>&1; echo "ncu list rc=$?"; tail -3 gpurun_out/profile_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gGemmTcgen05 -s 700 -c 6 -o gpurun_out/prof_gemm -f python scripts/profile_step.py 1 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"; tail -2 gpurun_out/ncu_gemm.log
fi
