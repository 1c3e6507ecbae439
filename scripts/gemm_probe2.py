"""Standalone vs in-step gap of the big products: the vocabulary projection (3200 x 32000 x 512) and the feed-forward up
projection with / without bias, warm (rotating buffers) and after an L2 flush (a 512 MB memset between launches)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
st = torch.cuda.Stream(); torch.cuda.set_stream(st); lib.set_stream(st.cuda_stream)
g = lib.gemm(4)
rs = np.random.RandomState(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def span(fn, cold):
    stamps = torch.zeros(64 * 148, dtype=torch.int64, device="cuda")
    res = []
    for it in range(6):
        if cold:
            flush.zero_()
        torch.cuda.synchronize()
        stamps.zero_(); torch.cuda.synchronize()
        lib.call("mrn_gemm_debug_stamps", stamps.data_ptr())
        fn(it)
        lib.call("mrn_gemm_debug_stamps", None)
        torch.cuda.synchronize()
        s = stamps.cpu().numpy().reshape(-1, 64)
        s = s[s[:, 0] > 0]
        if len(s):
            res.append((s[:, 63].max() - s[:, 0].min()) / 1000.0)
    return np.median(res[1:]) if len(res) > 1 else float("nan")
for (M, K, N, label) in [(3200, 512, 32000, "logits fwd"), (3200, 512, 2048, "ffn up")]:
    n = 3
    A = [lib.array(0.5 * rs.standard_normal((M, K)).astype(np.float32)) for _ in range(n)]
    B = [lib.array(0.05 * rs.standard_normal((K, N)).astype(np.float32)) for _ in range(n)]
    bias = lib.array(rs.standard_normal((1, N)).astype(np.float32))
    C = [lib.zeros((M, N)) for _ in range(n)]
    for cold in (False, True):
        a = span(lambda i: lib.call("mrn_prod", g.h, C[i % n].t(), A[i % n].t(), B[i % n].t(), 0, 0, 0.0, 1.0), cold)
        b = span(lambda i: lib.call("mrn_prod_affine", g.h, C[i % n].t(), A[i % n].t(), B[i % n].t(), bias.t()), cold)
        print("%s  %s: prod %.1f us, affine %.1f us (kernel spans, persistent kernel)" % (label, "after L2 flush" if cold else "warm", a, b), flush=True)
