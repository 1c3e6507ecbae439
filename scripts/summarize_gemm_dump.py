"""Per-shape table of the tensor-core GEMM launches of ONE replayed training step.

Input: the CSV written by `MRN_GEMM_PROFILE_DUMP=<file> python bench.py` (csrc/kernels/gemm.cu,
ProfileScope): CUDA events recorded as external event nodes around every launch INSIDE the
captured step graph, so durations include the in-graph scheduling gap of that launch and overlap
with side-stream work.  Columns: M,N,K,batches,layout(A,B: N = as stored, T = transposed),BN,splits,beta,us.

  python scripts/summarize_gemm_dump.py gpurun_out/gemm_dump_side.csv > profiles/gemm_in_graph_rNN.md
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 1458.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass

agg = collections.OrderedDict()
tot_us = 0.0
tot_flop = 0.0
for r in csv.reader(open(sys.argv[1])):
    key = tuple(r[:-1])
    us = float(r[-1])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
    tot_us += us
    tot_flop += 2.0 * int(r[0]) * int(r[1]) * int(r[2]) * int(r[3])
print("| M | N | K | op(A) op(B) | tile N | split-K | beta | launches/step | mean us | us/step | TFLOP/s | frac of %.0f TF/s |" % peak)
print("|---:|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, K, b = int(k[0]), int(k[1]), int(k[2]), int(k[3])
    tf = 2.0 * M * N * K * b / (us / c) / 1e6
    print("| %d | %d | %d | %s | %s | %s | %s | %d | %.1f | %.1f | %.1f | %.3f |" % (M, N, K, k[4], k[5], k[6], k[7][:3], c, us / c, us, tf, tf / peak))
print("\nall launches: %d per step, %.3f ms per step, %.1f GFLOP per step -> %.1f TFLOP/s = %.3f of the measured dense bf16 peak"
      % (sum(c for c, _ in agg.values()), tot_us / 1e3, tot_flop / 1e9, tot_flop / tot_us / 1e6, tot_flop / tot_us / 1e6 / peak))
