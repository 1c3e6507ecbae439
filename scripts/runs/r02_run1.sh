#!/bin/bash
# round 2, GPU call 1: full test suite, headline bench + CPU arm, config C bench + launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r1_gpu.txt
nproc > gpurun_out/r1_nproc.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r1_nproc.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
( time timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r1_bench.log 2>&1
( time timeout 300 python bench.py --impl reference --gpus 1 --steps 30 --warmup 5 ) > gpurun_out/r1_bench_ref.log 2>&1
( time timeout 600 python bench.py --model s2s-deep-gru --steps 10 --warmup 3 ) > gpurun_out/r1_bench_gru.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/r1_gru_launches.csv \
   python bench.py --model s2s-deep-gru --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r1_gru_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r1_gru_launches.csv > gpurun_out/r1_gru_launches_summary.txt 2>&1
tail -3 gpurun_out/r1_pytest.log; tail -2 gpurun_out/r1_bench.log; tail -2 gpurun_out/r1_bench_ref.log; tail -2 gpurun_out/r1_bench_gru.log
