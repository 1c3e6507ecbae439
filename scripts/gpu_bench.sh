#!/bin/bash
# bench line + ncu launch list + one full capture of the tensor-core GEMM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 2 1 > gpurun_out/profile_step.log 2>&1; echo "ncu list rc=$?"; tail -3 gpurun_out/profile_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gGemmTcgen05 -s 40 -c 4 -o gpurun_out/prof_gemm -f python scripts/profile_step.py 1 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out | head -30
