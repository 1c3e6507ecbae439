"""GEMM timing probe: one shape, beta 0/1, modes, with CUDA events (engine on a torch side stream)."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
lib.set_stream(s.cuda_stream)
def t(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters): fn()
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000  # us
rs = np.random.RandomState(0)
shapes = [(3200, 512, 512), (3200, 512, 2048), (3200, 2048, 512), (3200, 512, 32000), (512, 3200, 512), (512, 3200, 32000), (3200, 32000, 512)]
quick = len(sys.argv) > 1
for (M, K, N) in (shapes[:2] if quick else shapes):
    A = lib.array(rs.standard_normal((M, K)).astype(np.float32)); B = lib.array(rs.standard_normal((K, N)).astype(np.float32)); C = lib.zeros((M, N))
    for mode in (1,) if quick else (1, 2):
        g = lib.gemm(mode)
        for beta in (0.0, 1.0):
            us = t(lambda: lib.call("mrn_prod", g.h, C.t(), A.t(), B.t(), 0, 0, beta, 1.0))
            print(json.dumps({"M": M, "K": K, "N": N, "mode": mode, "beta": beta, "us_incl_pack": us, "tflops": 2.0 * M * N * K / us / 1e6}), flush=True)
    del A, B, C
