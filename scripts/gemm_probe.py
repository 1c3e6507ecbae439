"""GEMM tuning probe (tf32 mode): the step's product shapes under forced tile width / split-K
(MRN_GEMM_BN, MRN_GEMM_SPLITS), timed back to back with CUDA events on rotating buffers."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
s = torch.cuda.Stream(); torch.cuda.set_stream(s); lib.set_stream(s.cuda_stream)
g = lib.gemm(3)
rs = np.random.RandomState(0)
def t(fn, n, iters=40, warm=5):
    for i in range(warm): fn(i % n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for i in range(iters): fn(i % n)
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000
# (rowsA, colsA, rowsB, colsB, tA, tB, beta, label)
R, D, F = 3200, 512, 2048
cases = [(R, D, D, D, 0, 0, 0.0, "proj fwd"), (R, D, D, D, 0, 1, 1.0, "proj dX"), (R, D, R, D, 1, 0, 1.0, "proj dW"),
         (R, D, D, F, 0, 0, 0.0, "ffn1 fwd"), (R, F, F, D, 0, 0, 0.0, "ffn2 fwd"), (R, D, F, D, 0, 1, 0.0, "ffn2 dX"), (R, F, D, F, 0, 1, 1.0, "ffn1 dX"),
         (R, D, R, F, 1, 0, 1.0, "ffn1 dW"), (R, F, R, D, 1, 0, 1.0, "ffn2 dW")]
for (ra, ca, rb, cb, tA, tB, beta, label) in cases:
    M = ca if tA else ra; K = ra if tA else ca; N = rb if tB else cb
    n = 4
    A = [lib.array(rs.standard_normal((ra, ca)).astype(np.float32)) for _ in range(n)]
    B = [lib.array(rs.standard_normal((rb, cb)).astype(np.float32)) for _ in range(n)]
    C = [lib.zeros((M, N)) for _ in range(n)]
    row = {"case": label, "M": M, "N": N, "K": K}
    for bn in ("64", "128"):
        for sp in ("auto", "1", "2", "4"):
            os.environ["MRN_GEMM_BN"] = bn
            if sp == "auto": os.environ.pop("MRN_GEMM_SPLITS", None)
            else: os.environ["MRN_GEMM_SPLITS"] = sp
            us = t(lambda i: lib.call("mrn_prod", g.h, C[i].t(), A[i].t(), B[i].t(), tA, tB, beta, 1.0), n)
            row["bn%s_sp%s" % (bn, sp)] = round(us, 1)
    print(json.dumps(row), flush=True)
    del A, B, C
