"""Times the fused attention kernels at the config-B head shape (CUDA events, rotating buffers)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
side = torch.cuda.Stream(); lib.set_stream(side.cuda_stream)
B, H, T, dk = 64, 8, 50, 64; D = H * dk
rnd = lambda *s: np.random.standard_normal(s).astype(np.float32)
def mk():
    return (lib.array(rnd(B, T, D)), lib.array(rnd(B, T, D)), lib.array(rnd(B, T, D)), lib.zeros((B, T, D)), lib.zeros((B, H, T, T)),
            lib.array(rnd(B, T, D)), lib.zeros((B, T, D)), lib.zeros((B, T, D)), lib.zeros((B, T, D)))
pool = [mk() for _ in range(6)]
mask = lib.array(np.zeros((B, 1, 1, T), np.float32))
def timeit(fn, iters=50, warm=5):
    # the launches are captured into one CUDA graph and the replay is timed (the ctypes call path
    # costs more than the forward kernel)
    for i in range(warm): fn(pool[i % len(pool)])
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
        for i in range(iters): fn(pool[i % len(pool)])
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record(side); graph.replay(); e1.record(side)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000
for exact in (0, 1):
    f = timeit(lambda s: lib.call("mrn_multi_head_attention", s[3].t(), s[4].t(), s[0].t(), s[1].t(), s[2].t(), mask.t(), H, 0.125, exact))
    b = timeit(lambda s: lib.call("mrn_multi_head_attention_grad", s[6].t(), s[7].t(), s[8].t(), s[5].t(), s[3].t(), s[4].t(), s[0].t(), s[1].t(), s[2].t(), H, 0.125, exact))
    print(json.dumps({"simt": bool(os.environ.get("MRN_ATTENTION_SIMT")), "exact": exact, "fwd_us": f, "bwd_us": b}))
