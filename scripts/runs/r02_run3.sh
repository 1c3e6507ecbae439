#!/bin/bash
# round 2, GPU call 3: where does the bf16 step spend its time (launch list) + ncu --set full of three product shapes
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1400 --csv --log-file gpurun_out/r3_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-graph-replay > gpurun_out/r3_launch_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r3_launches.csv > gpurun_out/r3_launches_summary.txt 2>&1
for spec in "logits_fwd 4 3200 512 512 32000 0 0 0.0 6" "ffn_up 4 3200 512 512 2048 0 0 0.0 10" "proj_fwd 4 3200 512 512 512 0 0 0.0 10" "proj_dw 4 3200 512 3200 512 1 0 1.0 10" "ffn_gated 4 3200 512 2048 512 0 1 0.0 10 gate"; do
  set -- $spec; name=$1; shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gGemmBf16 -s 4 -c 1 -o gpurun_out/r3_ncu_$name python scripts/gemm_one.py "$@" > gpurun_out/r3_ncu_$name.log 2>&1
  timeout 120 python scripts/gemm_one.py "$@" >> gpurun_out/r3_times.txt 2>&1
done
head -30 gpurun_out/r3_launches_summary.txt; cat gpurun_out/r3_times.txt
