"""One product shape, launched repeatedly on rotating buffers - the target of `ncu --set full` captures
and of quick A/B timings.   python scripts/gemm_one.py MODE rowsA colsA rowsB colsB tA tB beta [iters] [gate]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
s = torch.cuda.Stream(); torch.cuda.set_stream(s); lib.set_stream(s.cuda_stream)
mode, ra, ca, rb, cb, tA, tB = [int(x) for x in sys.argv[1:8]]
beta = float(sys.argv[8]); iters = int(sys.argv[9]) if len(sys.argv) > 9 else 20
gate = len(sys.argv) > 10 and sys.argv[10] == "gate"
g = lib.gemm(mode)
rs = np.random.RandomState(0)
M = ca if tA else ra; K = ra if tA else ca; N = rb if tB else cb
n = 3
A = [lib.array(rs.standard_normal((ra, ca)).astype(np.float32)) for _ in range(n)]
B = [lib.array(rs.standard_normal((rb, cb)).astype(np.float32)) for _ in range(n)]
C = [lib.zeros((M, N)) for _ in range(n)]
H = [lib.array(rs.standard_normal((M, N)).astype(np.float32)) for _ in range(n)] if gate else None
def run(i):
    if gate:
        lib.call("mrn_prod_swish_grad_nt", g.h, C[i].t(), A[i].t(), B[i].t(), H[i].t(), beta)
    else:
        lib.call("mrn_prod", g.h, C[i].t(), A[i].t(), B[i].t(), tA, tB, beta, 1.0)
for i in range(3): run(i % n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for i in range(iters): run(i % n)
e1.record(s); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / iters * 1000
print("mode %d  M=%d N=%d K=%d tA=%d tB=%d beta=%g%s: %.1f us per call (incl. operand conversion in mode 4), %.1f TFLOP/s" % (mode, M, N, K, tA, tB, beta, " gated" if gate else "", us, 2.0 * M * N * K / us / 1e6))
