#!/bin/bash
# round 2, GPU call 39: the per-GPU share of config E (Transformer-big, 32 x 80 of the 256 x 80 global batch) on one B200, final code
mkdir -p gpurun_out
timeout 150 python bench.py --model transformer-big --batch 32 --len 80 --steps 10 --warmup 3 --no-cpu-baseline --no-traffic --no-parity > gpurun_out/r39_big.json 2> gpurun_out/r39_big.err; echo "rc=$?"; cut -c1-420 gpurun_out/r39_big.json; tail -3 gpurun_out/r39_big.err
