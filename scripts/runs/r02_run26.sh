#!/bin/bash
mkdir -p gpurun_out
for t in "3200,512,2048,1.0,128,1" "3200,512,2048,1.0,128,2" "3200,512,1536,1.0,128,1" "512,2048,3200,1.0,128,2" "2048,512,3200,1.0,64,3"; do
  echo "== TRY $t"; MRN_GEMM_TRY=$t timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity 2>&1 | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['gemm_ms_per_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 1100 --csv --log-file gpurun_out/r26_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --no-graph-replay > gpurun_out/r26_launch_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r26_launches.csv > gpurun_out/r26_launches_summary.txt 2>&1
for spec in "ffn_up 4 3200 512 512 2048 0 0 0.0 10" "logits_fwd 4 3200 512 512 32000 0 0 0.0 6" "ffn_gated 4 3200 512 2048 512 0 1 0.0 10 gate"; do
  set -- $spec; name=$1; shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gGemmBf16 -s 4 -c 1 -o gpurun_out/r26_ncu_$name python scripts/gemm_one.py "$@" > gpurun_out/r26_ncu_$name.log 2>&1
done
head -24 gpurun_out/r26_launches_summary.txt
