"""Per-operator roofline table at the config-B shapes: this repo's sm_100a kernels next to
the reference's own kernels (oracle/_ref, recompiled for sm_100a), CUDA-event timed.

  python scripts/op_bench.py > gpurun_out/op_table.json     (GPU box)

Algorithmic bytes follow SURVEY.md 8d ("read each input once, write each output once,
+1 read when the op accumulates into its output"); GEMM: 2MNK flops.  Buffers rotate
through a pool larger than the 126 MB L2 so every launch reads from HBM.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft

pkg = graft.load_package()
lib = pkg.load()
# our kernels run on a side stream, the reference's kernels on the legacy default stream;
# timeit() records its events on whichever stream the timed callable uses
side = torch.cuda.Stream()
lib.set_stream(side.cuda_stream)
REF = os.path.join(ROOT, "oracle", "_ref", "libmarian_ref_kernels.so")
ref = ctypes.CDLL(REF) if os.path.exists(REF) else None
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
F = ctypes.c_float


def rnd(*shape, scale=1.0):
    return (np.random.standard_normal(shape) * scale).astype(np.float32)


def timeit(fn, pool, iters=30, warm=5, stream=None):
    """ms per call.  Our kernels (side stream) are captured into ONE CUDA graph of `iters` launches and
    the replay is timed: the python/ctypes call path costs 15-25 us per call, more than most of the
    kernels measured here.  The reference's kernels run on the legacy stream (not capturable) through a
    plain C call and stay on the eager loop."""
    own = stream is None
    stream = side if own else stream
    for i in range(warm):
        fn(pool[i % len(pool)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph = None
    if own and not os.environ.get("OPBENCH_EAGER"):
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream, capture_error_mode="thread_local"):
                for i in range(iters):
                    fn(pool[i % len(pool)])
        except Exception as e:  # an op that cannot be captured falls back to the eager loop
            print("capture failed, eager timing: %s" % e, file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    if graph is not None:
        graph.replay()  # warm replay (uploads the graph)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            e0.record(stream)
            graph.replay()
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    e0.record(stream)
    for i in range(iters):
        fn(pool[i % len(pool)])
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


LEGACY = torch.cuda.default_stream()


def pool_of(make, bytes_per_set):
    n = max(2, int(300e6 // max(bytes_per_set, 1)) + 1)
    return [make() for _ in range(min(n, 24))]


rows = []


def report(name, alg_bytes, mine_ms, ref_ms=None, flops=None):
    r = {"op": name, "ms": mine_ms}
    if flops:
        r["tflops"] = flops / (mine_ms * 1e-3) / 1e12
        r["frac_of_bf16_peak"] = r["tflops"] / peaks["bf16_tflops"]
        if ref_ms:
            r["ref_tflops"] = flops / (ref_ms * 1e-3) / 1e12
    else:
        r["gbs"] = alg_bytes / (mine_ms * 1e-3) / 1e9
        r["frac_of_hbm_peak"] = r["gbs"] / peaks["hbm_gbs"]
        if ref_ms:
            r["ref_gbs"] = alg_bytes / (ref_ms * 1e-3) / 1e9
    if ref_ms:
        r["ref_ms"] = ref_ms
        r["speedup_vs_ref_kernel"] = ref_ms / mine_ms
    rows.append(r)
    print(json.dumps(r), flush=True)


R, D, FF, V = 3200, 512, 2048, 32000

# ---- layer norm fwd / bwd --------------------------------------------------
gamma, beta = lib.array(1 + 0.1 * rnd(1, D)), lib.array(0.1 * rnd(1, D))
bt = beta.t()
pool = pool_of(lambda: (lib.array(rnd(R, D)), lib.zeros((R, D)), lib.array(rnd(R, D)), lib.zeros((R, D)), lib.zeros((1, D)), lib.zeros((1, D))), 4 * R * D * 4)
mine = timeit(lambda s: lib.call("mrn_layer_norm", s[1].t(), s[0].t(), gamma.t(), bt, 1e-6), pool)
theirs = timeit(lambda s: ref.ref_layer_norm(s[1].t(), s[0].t(), gamma.t(), ctypes.byref(bt), F(1e-6)), pool, stream=LEGACY) if ref else None
report("LayerNormalization 3200x512", 2 * R * D * 4, mine, theirs)
mine = timeit(lambda s: lib.call("mrn_layer_norm_grad", s[3].t(), s[4].t(), s[5].t(), s[2].t(), s[1].t(), s[0].t(), gamma.t(), bt, 1e-6), pool)
theirs = timeit(lambda s: ref.ref_layer_norm_grad(s[3].t(), s[4].t(), ctypes.byref(s[5].t()), s[2].t(), s[1].t(), s[0].t(), gamma.t(), ctypes.byref(bt), F(1e-6)), pool, stream=LEGACY) if ref else None
report("LayerNormalizationGrad 3200x512", 5 * R * D * 4, mine, theirs)

# ---- softmax fwd / bwd (attention scores) ------------------------------------
S = (64, 8, 50, 50)
n = int(np.prod(S))
pool = pool_of(lambda: (lib.array(rnd(*S)), lib.zeros(S), lib.array(rnd(*S)), lib.zeros(S)), 4 * n * 4)
mine = timeit(lambda s: lib.call("mrn_softmax", s[1].t(), s[0].t(), None), pool)
theirs = timeit(lambda s: ref.ref_softmax(s[1].t(), s[0].t(), None), pool, stream=LEGACY) if ref else None
report("Softmax 25600x50", 2 * n * 4, mine, theirs)
mine = timeit(lambda s: lib.call("mrn_softmax_grad", s[3].t(), s[2].t(), s[1].t()), pool)
theirs = timeit(lambda s: ref.ref_softmax_grad(s[3].t(), s[2].t(), s[1].t()), pool, stream=LEGACY) if ref else None
report("SoftmaxGrad 25600x50", 4 * n * 4, mine, theirs)

# ---- fused multi-head attention (one encoder self-attention block of config B) ------
Bq, Hh, Tt, dk = 64, 8, 50, 64
def _attn_set():
    return (lib.array(rnd(Bq, Tt, D)), lib.array(rnd(Bq, Tt, D)), lib.array(rnd(Bq, Tt, D)), lib.zeros((Bq, Tt, D)), lib.zeros((Bq, Hh, Tt, Tt)),
            lib.array(rnd(Bq, Tt, D)), lib.zeros((Bq, Tt, D)), lib.zeros((Bq, Tt, D)), lib.zeros((Bq, Tt, D)))
amask = lib.array(np.zeros((Bq, 1, 1, Tt), dtype=np.float32))
pool = pool_of(_attn_set, 9 * R * D * 4)
mine = timeit(lambda s: lib.call("mrn_multi_head_attention", s[3].t(), s[4].t(), s[0].t(), s[1].t(), s[2].t(), amask.t(), Hh, 0.125, 0), pool)
# algorithmic bytes: q, k, v in; out and probs out
report("MultiHeadAttention fwd B64 H8 T50 dk64 (fused; reference = 9 kernels)", (4 * R * D + Bq * Hh * Tt * Tt) * 4, mine, None)
mine = timeit(lambda s: lib.call("mrn_multi_head_attention_grad", s[6].t(), s[7].t(), s[8].t(), s[5].t(), s[3].t(), s[4].t(), s[0].t(), s[1].t(), s[2].t(), Hh, 0.125, 0), pool)
# q, k, v, out, dout, probs in; dq, dk, dv read + written
report("MultiHeadAttentionGrad B64 H8 T50 dk64 (fused)", (11 * R * D + Bq * Hh * Tt * Tt) * 4, mine, None)
del pool

# ---- cross entropy fwd / bwd at the logits shape -------------------------------
pick = lib.array(np.random.randint(0, V, size=(R, 1)).astype(np.float32))
adj = lib.array(rnd(R, 1))
pool = [(lib.array(rnd(R, V)), lib.zeros((R, 1)), lib.zeros((R, V))) for _ in range(2)]
mine = timeit(lambda s: lib.call("mrn_cross_entropy_pick", s[1].t(), s[0].t(), pick.t()), pool, iters=10, warm=2)
theirs = timeit(lambda s: ref.ref_cross_entropy_pick(s[1].t(), s[0].t(), pick.t()), pool, iters=10, warm=2, stream=LEGACY) if ref else None
report("CrossEntropyPick 3200x32000", R * V * 4, mine, theirs)
mine = timeit(lambda s: lib.call("mrn_cross_entropy_pick_backward", s[2].t(), adj.t(), s[0].t(), pick.t()), pool, iters=10, warm=2)
theirs = timeit(lambda s: ref.ref_cross_entropy_pick_backward(s[2].t(), adj.t(), s[0].t(), pick.t()), pool, iters=10, warm=2, stream=LEGACY) if ref else None
report("CrossEntropyPickBackward 3200x32000", 3 * R * V * 4, mine, theirs)
del pool

# ---- element-wise: residual add, swish, bias gradient -------------------------
pool = pool_of(lambda: (lib.array(rnd(R, D)), lib.array(rnd(R, D)), lib.zeros((R, D))), 3 * R * D * 4)
mine = timeit(lambda s: lib.call("mrn_element", b"plus", s[2].t(), lib.tensor_list([s[0].t(), s[1].t()]), 2, 0.0), pool)
theirs = timeit(lambda s: ref.ref_element(b"plus", s[2].t(), lib.tensor_list([s[0].t(), s[1].t()]), 2, F(0)), pool, stream=LEGACY) if ref else None
report("Element plus 3200x512", 3 * R * D * 4, mine, theirs)
pool = pool_of(lambda: (lib.array(rnd(R, FF)), lib.zeros((R, FF))), 2 * R * FF * 4)
mine = timeit(lambda s: lib.call("mrn_element", b"swish", s[1].t(), lib.tensor_list([s[0].t()]), 1, 0.0), pool)
theirs = timeit(lambda s: ref.ref_element(b"swish", s[1].t(), lib.tensor_list([s[0].t()]), 1, F(0)), pool, stream=LEGACY) if ref else None
report("Element swish 3200x2048", 2 * R * FF * 4, mine, theirs)
bias_out = lib.zeros((1, FF))
mine = timeit(lambda s: lib.call("mrn_add", b"id", 1.0, bias_out.t(), lib.tensor_list([s[0].t()]), 1, 0.0), pool)
theirs = timeit(lambda s: ref.ref_add(b"id", F(1), bias_out.t(), lib.tensor_list([s[0].t()]), 1, F(0)), pool, iters=5, warm=1, stream=LEGACY) if ref else None
report("Add bias-gradient [1,2048] <- [3200,2048]", R * FF * 4, mine, theirs)

# ---- transposes -------------------------------------------------------------
axes = (ctypes.c_int * 4)(0, 2, 1, 3)
pool = pool_of(lambda: (lib.array(rnd(64, 50, 8, 64)), lib.zeros((64, 8, 50, 64))), 2 * R * D * 4)
mine = timeit(lambda s: lib.call("mrn_transpose_nd", s[1].t(), s[0].t(), axes), pool)
theirs = timeit(lambda s: ref.ref_transpose_nd(s[1].t(), s[0].t(), axes), pool, stream=LEGACY) if ref else None
report("TransposeND {0,2,1,3} 64x50x8x64", 2 * R * D * 4, mine, theirs)

# ---- Adam over the 93M-parameter arena ----------------------------------------
P = 93_326_336
p, g, m, v = (lib.array(rnd(1, P, scale=0.01)) for _ in range(4))
one = [(0,)]
mine = timeit(lambda s: lib.call("mrn_adam_step", p.t(), g.t(), m.t(), v.t(), 1e-4, 0.9, 0.999, 1e-8, 1, 1.0, 1.0), one, iters=10, warm=2)
theirs = timeit(lambda s: ref.ref_adam_step(p.t(), g.t(), m.t(), v.t(), F(1e-4), F(0.9), F(0.999), F(1e-8), 1, F(1.0)), one, iters=5, warm=1, stream=LEGACY) if ref else None
report("clip + Adam, P = 93.3M", 8 * P * 4, mine, theirs)
del p, g, m, v

# ---- GEMMs: tcgen05 bf16 / bf16x3 vs the reference's cublasSgemm ----------------
for (M, K, N, tA, tB, name) in ((R, D, D, 0, 0, "proj fwd"), (R, D, FF, 0, 0, "ffn-up fwd"), (R, FF, D, 0, 0, "ffn-down fwd"), (R, D, V, 0, 0, "logits fwd"),
                                (R, V, D, 0, 1, "logits dX"), (D, R, V, 1, 0, "logits dW (A^T)"), (D, R, D, 1, 0, "proj dW (A^T, split-K)")):
    a_shape = (K, M) if tA else (M, K)
    b_shape = (N, K) if tB else (K, N)
    A, B, C = lib.array(rnd(*a_shape)), lib.array(rnd(*b_shape)), lib.zeros((M, N))
    flops = 2.0 * M * N * K
    theirs = timeit(lambda s: ref.ref_prod(C.t(), A.t(), B.t(), tA, tB, F(0), F(1)), one, iters=5, warm=2, stream=LEGACY) if ref else None
    for mode, mname in ((3, "tf32 (no packing)"), (1, "bf16 (incl. operand packing)"), (2, "bf16x3 (incl. operand packing)")):
        gm = lib.gemm(mode)
        mine = timeit(lambda s: lib.call("mrn_prod", gm.h, C.t(), A.t(), B.t(), tA, tB, 0.0, 1.0), one, iters=10, warm=3)
        report("Prod %s %dx%dx%d %s" % (name, M, N, K, mname), 0, mine, theirs, flops=flops)
    del A, B, C

print(json.dumps({"peaks": peaks, "rows": rows}))
