#!/bin/bash
# bench line (+ optional op table, ncu launch list and ncu --set full captures of the top kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MODE=${MODE:-3}
if [ -z "$NOBENCH" ]; then
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [ -n "$OPBENCH" ]; then timeout 900 python scripts/op_bench.py > gpurun_out/op_table.json 2> gpurun_out/op_table.err; echo "op_bench rc=$?"; tail -3 gpurun_out/op_table.err; fi
if [ -n "$NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 2 $MODE > gpurun_out/profile_step.log 2>&1; echo "ncu list rc=$?"; tail -3 gpurun_out/profile_step.log
fi
if [ -n "$NCUFULL" ]; then
# one full capture per entry of $NCUFULL: "kernelRegex[:skip[:count]]" (space separated)
for spec in $NCUFULL; do
  k=${spec%%:*}; rest=${spec#*:}; skip=${NCUSKIP:-100}; cnt=${NCUCOUNT:-4}
  if [ "$rest" != "$spec" ]; then skip=${rest%%:*}; r2=${rest#*:}; if [ "$r2" != "$rest" ]; then cnt=$r2; fi; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -o gpurun_out/prof_$k -f python scripts/profile_step.py 2 $MODE > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k rc=$?"; tail -2 gpurun_out/ncu_$k.log
done
fi
if [ -n "$NCUDRAM" ]; then
# DRAM bytes + duration of every tensor-core GEMM launch of one eager step (roofline.traffic)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gGemm -c 2000 --csv --log-file gpurun_out/gemm_dram.csv python scripts/profile_step.py 2 $MODE > gpurun_out/ncu_dram.log 2>&1; echo "ncu dram rc=$?"
fi
