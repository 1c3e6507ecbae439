"""Parity of every CUDA tensor operator with the CPU oracle, through the C ABI,
on the same seeded inputs.  Shapes include the ragged / unaligned / broadcast
cases the reference's kernels have code paths for, plus the config-B shapes.
Tolerance: 1e-5 relative to the tensor's magnitude for fp32 element-wise/row
ops (summation order and libm-vs-CUDA expf differ in the last bits)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 2e-5


def close(a, b, rtol=RTOL, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(1e-6, float(np.abs(b).max()))
    err = float(np.abs(a - b).max()) / scale
    assert err <= rtol, "%s: max err %.3e (scale %.3e)" % (what, err, scale)


def both(cuda, oracle, fn):
    """Runs fn(lib) on both libraries and compares every returned array."""
    got = fn(cuda)
    exp = fn(oracle)
    cuda.synchronize()
    assert got.keys() == exp.keys()
    return got, exp


def compare(cuda, oracle, fn, rtol=RTOL):
    got, exp = both(cuda, oracle, fn)
    for k in exp:
        close(got[k], exp[k], rtol, k)


def rnd(seed, *shape, scale=1.0):
    return (np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32)


# ---------------------------------------------------------------- element-wise
@pytest.mark.parametrize("functor,nin", [("plus", 2), ("minus", 2), ("mult", 2), ("div", 2), ("tanh3", 3), ("swish", 1),
                                         ("logit", 1), ("relu", 1), ("scale", 1), ("shift", 1), ("neg", 1), ("exp", 1), ("square", 1)])
@pytest.mark.parametrize("shape", [(3, 50, 64), (7, 13), (4, 1000)])
def test_element_same_shape(cuda, oracle, functor, nin, shape):
    def fn(lib):
        ins = [lib.array(rnd(10 + i, *shape) + (2.5 if functor == "div" and i == 1 else 0)) for i in range(nin)]
        out = lib.zeros(shape)
        lib.call("mrn_element", functor.encode(), out.t(), lib.tensor_list([x.t() for x in ins]), nin, 0.37)
        return {"out": out.numpy()}

    compare(cuda, oracle, fn)


@pytest.mark.parametrize("shapes", [
    ((2, 50, 64, 512 // 8), (50, 1, 64)),          # positional signal [T,1,d] onto [.,T,B,d]
    ((4, 8, 10, 10), (4, 1, 1, 10)),               # attention mask [B,1,1,T]
    ((4, 8, 10, 10), (4, 1, 10, 10)),              # decoder self mask
    ((6, 5, 3), (6, 5, 1)),                        # [T,B,d] * mask[T,B,1]
    ((33, 7), (1, 7)),                             # row vector
    ((33, 7), (33, 1)),                            # column vector
    ((2, 2, 1), (2, 1)),                           # the reference's own broadcast test
])
def test_element_broadcast(cuda, oracle, shapes):
    full, small = shapes

    def fn(lib):
        a, b = lib.array(rnd(1, *full)), lib.array(rnd(2, *small))
        out1, out2 = lib.zeros(full), lib.zeros(full)
        lib.call("mrn_element", b"plus", out1.t(), lib.tensor_list([a.t(), b.t()]), 2, 0.0)
        lib.call("mrn_element", b"mult", out2.t(), lib.tensor_list([b.t(), a.t()]), 2, 0.0)
        return {"plus": out1.numpy(), "mult": out2.numpy()}

    compare(cuda, oracle, fn)


def test_element_inplace(cuda, oracle):
    def fn(lib):
        a, o = lib.array(rnd(1, 1000)), lib.array(rnd(2, 1000))
        lib.call("mrn_element", b"axpy_self", o.t(), lib.tensor_list([a.t()]), 1, -0.5)
        return {"o": o.numpy()}

    compare(cuda, oracle, fn)


# ---------------------------------------------------------------- Add: 3 cases
@pytest.mark.parametrize("out_shape,in_shape", [
    ((1, 512), (1, 50, 64, 512)),     # bias gradient: column sums            (generic, case 3)
    ((1, 7), (33, 7)),
    ((1, 1, 64, 1), (1, 50, 64, 1)),  # cost: sum over time                   (generic)
    ((1, 50, 64, 1), (1, 50, 64, 37)),  # last-axis reduction                 (case 1)
    ((5, 1), (5, 1000)),
    ((50, 1, 16), (50, 64, 16)),      # reduce a middle axis
    ((1, 64, 2048), (50, 64, 2048)),  # leading axis reduced, two trailing axes kept (case 3b)
    ((1, 1, 5, 36), (3, 7, 5, 36)),
    ((1, 5, 35), (7, 5, 35)),         # same shape class, unaligned rows: generic kernel
    ((3, 50, 16), (3, 50, 16)),       # plain accumulate                      (case 2)
    ((50, 64, 16), (50, 64, 1)),      # broadcast input accumulated into full (case 2 + bcast)
])
def test_add_reduce(cuda, oracle, out_shape, in_shape):
    def fn(lib):
        a = lib.array(rnd(3, *in_shape))
        out = lib.array(rnd(4, *out_shape))
        lib.call("mrn_add", b"id", 0.5, out.t(), lib.tensor_list([a.t()]), 1, 0.0)
        return {"out": out.numpy()}

    compare(cuda, oracle, fn, rtol=5e-5)


def test_add_two_inputs_broadcast_reduce(cuda, oracle):
    # scalar_product backward pattern: grad[T,B,D] += ctx-like * adj[1,B,D]
    def fn(lib):
        a, b = lib.array(rnd(5, 20, 8, 32)), lib.array(rnd(6, 1, 8, 32))
        o1, o2 = lib.array(rnd(7, 20, 8, 32)), lib.array(rnd(8, 1, 8, 32))
        lib.call("mrn_add", b"mult", 1.0, o1.t(), lib.tensor_list([a.t(), b.t()]), 2, 0.0)
        lib.call("mrn_add", b"mult", 1.0, o2.t(), lib.tensor_list([a.t(), b.t()]), 2, 0.0)   # reduces over axis 0
        lib.call("mrn_add", b"tanh_grad", 1.0, o1.t(), lib.tensor_list([a.t(), a.t()]), 2, 0.0)
        return {"o1": o1.numpy(), "o2": o2.numpy()}

    compare(cuda, oracle, fn, rtol=5e-5)


def test_reduce_scalar_product_pattern(cuda, oracle):
    # scalar_product forward of the recurrent attention: out[1,B,C] = sum_t ctx[t,b,c] * e[t,b]   (case 3b, broadcast operand)
    def fn(lib):
        ctx, e = lib.array(rnd(5, 50, 64, 512)), lib.array(rnd(6, 50, 64, 1))
        o1 = lib.array(rnd(7, 1, 64, 512))
        lib.call("mrn_add", b"mult", 1.0, o1.t(), lib.tensor_list([ctx.t(), e.t()]), 2, 0.0)
        return {"o1": o1.numpy()}

    compare(cuda, oracle, fn, rtol=5e-5)


# ---------------------------------------------------------------- softmax family
@pytest.mark.parametrize("rows,cols", [(25, 50), (3, 1), (64, 257), (5, 3000), (2, 8)])
def test_softmax_family(cuda, oracle, rows, cols):
    def fn(lib):
        x = lib.array(rnd(1, rows, cols, scale=3))
        sm, lsm = lib.zeros((rows, cols)), lib.zeros((rows, cols))
        lib.call("mrn_softmax", sm.t(), x.t(), None)
        lib.call("mrn_logsoftmax", lsm.t(), x.t())
        adj = lib.array(rnd(2, rows, cols))
        g1, g2 = lib.array(rnd(3, rows, cols)), lib.array(rnd(3, rows, cols))
        lib.call("mrn_softmax_grad", g1.t(), adj.t(), sm.t())
        lib.call("mrn_logsoftmax_grad", g2.t(), adj.t(), lsm.t())
        return {"sm": sm.numpy(), "lsm": lsm.numpy(), "g1": g1.numpy(), "g2": g2.numpy()}

    compare(cuda, oracle, fn)


def test_softmax_reference_extremes(cuda, oracle):
    # the reference's own softmax test input (operator_tests.cpp:52) incl. +-100 logits
    def fn(lib):
        x = lib.array(np.array([-.2, -.3, 4.5, 5.2, -10, 101.45, -100.05, 1.05e-5], dtype=np.float32).reshape(2, 2, 2))
        sm = lib.zeros((2, 2, 2))
        lib.call("mrn_softmax", sm.t(), x.t(), None)
        return {"sm": sm.numpy()}

    compare(cuda, oracle, fn)


@pytest.mark.parametrize("mask_shape", [(6, 1, 1, 10), (6, 4, 10, 10), (1, 1, 1, 10)])
def test_softmax_masked(cuda, oracle, mask_shape):
    shape = (6, 4, 10, 10)

    def fn(lib):
        x = lib.array(rnd(1, *shape, scale=2))
        m = (np.random.RandomState(9).rand(*mask_shape) > 0.3).astype(np.float32)
        m[..., 0] = 1  # at least one unmasked entry per row
        mask = lib.array(m)
        out = lib.zeros(shape)
        mt = mask.t()
        lib.call("mrn_softmax", out.t(), x.t(), mt)
        return {"out": out.numpy()}

    compare(cuda, oracle, fn)


# ---------------------------------------------------------------- cross entropy
@pytest.mark.parametrize("rows,cols", [(37, 100), (16, 32000), (5, 1001), (3, 4)])
def test_cross_entropy(cuda, oracle, rows, cols):
    def fn(lib):
        x = lib.array(rnd(1, rows, cols, scale=2))
        pick = lib.array(np.random.RandomState(2).randint(0, cols, size=(rows, 1)).astype(np.float32))
        out = lib.zeros((rows, 1))
        lib.call("mrn_cross_entropy_pick", out.t(), x.t(), pick.t())
        adj = lib.array(rnd(3, rows, 1))
        g = lib.array(rnd(4, rows, cols, scale=0.1))
        lib.call("mrn_cross_entropy_pick_backward", g.t(), adj.t(), x.t(), pick.t())
        return {"ce": out.numpy(), "grad": g.numpy()}

    compare(cuda, oracle, fn)


# ---------------------------------------------------------------- layer norm
@pytest.mark.parametrize("rows,cols,eps,with_beta", [(64, 512, 1e-6, True), (33, 1024, 1e-9, True), (7, 3072, 1e-9, False), (5, 10, 1e-5, True),
                                                     (3200, 512, 1e-6, True), (130, 260, 1e-6, False), (50, 768, 1e-6, True)])
def test_layer_norm(cuda, oracle, rows, cols, eps, with_beta):
    def fn(lib):
        x = lib.array(rnd(1, rows, cols))
        gamma = lib.array(1 + 0.1 * rnd(2, 1, cols))
        beta = lib.array(0.1 * rnd(3, 1, cols)) if with_beta else None
        y = lib.zeros((rows, cols))
        bt = beta.t() if beta else None
        lib.call("mrn_layer_norm", y.t(), x.t(), gamma.t(), bt, eps)
        adj = lib.array(rnd(4, rows, cols))
        gx, gg = lib.array(rnd(5, rows, cols)), lib.array(rnd(6, 1, cols))
        gb = lib.array(rnd(7, 1, cols)) if with_beta else None
        gbt = gb.t() if gb else None
        lib.call("mrn_layer_norm_grad", gx.t(), gg.t(), gbt, adj.t(), y.t(), x.t(), gamma.t(), bt, eps)
        r = {"y": y.numpy(), "gx": gx.numpy(), "gg": gg.numpy()}
        if gb:
            r["gb"] = gb.numpy()
        return r

    compare(cuda, oracle, fn, rtol=5e-5)


@pytest.mark.parametrize("rows,cols", [(3200, 512), (77, 1024), (9, 36)])
def test_residual_layer_norm(cuda, oracle, rows, cols):
    """layer_norm(x + r) fused == the reference's Plus node followed by LayerNormalization (both
    libraries), forward and backward; the backward delivers the same gradient to x and r."""
    def fn(lib, fused):
        x, r = lib.array(rnd(1, rows, cols)), lib.array(rnd(2, rows, cols))
        gamma, beta = lib.array(1 + 0.1 * rnd(3, 1, cols)), lib.array(0.1 * rnd(4, 1, cols))
        adj = lib.array(rnd(5, rows, cols))
        y = lib.zeros((rows, cols))
        gx, gr = lib.array(rnd(6, rows, cols)), lib.array(rnd(7, rows, cols))
        gg, gb = lib.array(rnd(8, 1, cols)), lib.array(rnd(9, 1, cols))
        if fused:
            lib.call("mrn_residual_layer_norm", y.t(), x.t(), r.t(), gamma.t(), beta.t(), 1e-6)
            lib.call("mrn_residual_layer_norm_grad", gx.t(), gr.t(), gg.t(), gb.t(), adj.t(), y.t(), x.t(), r.t(), gamma.t(), beta.t(), 1e-6)
        else:
            s = lib.zeros((rows, cols))
            lib.call("mrn_element", b"plus", s.t(), lib.tensor_list([x.t(), r.t()]), 2, 0.0)
            lib.call("mrn_layer_norm", y.t(), s.t(), gamma.t(), beta.t(), 1e-6)
            gs = lib.zeros((rows, cols))
            lib.call("mrn_layer_norm_grad", gs.t(), gg.t(), gb.t(), adj.t(), y.t(), s.t(), gamma.t(), beta.t(), 1e-6)
            lib.call("mrn_add", b"id", 1.0, gx.t(), lib.tensor_list([gs.t()]), 1, 0.0)
            lib.call("mrn_add", b"id", 1.0, gr.t(), lib.tensor_list([gs.t()]), 1, 0.0)
        return {"y": y.numpy(), "gx": gx.numpy(), "gr": gr.numpy(), "gg": gg.numpy(), "gb": gb.numpy()}

    ref = fn(oracle, False)
    for lib, fused in ((cuda, True), (oracle, True), (cuda, False)):
        if lib is cuda and fused and cols == 36:
            pass  # 36 floats: still 16-byte aligned rows -> fused kernel applies
        got = fn(lib, fused)
        for k in ref:
            close(got[k], ref[k], 5e-5, "%s (%s fused=%s)" % (k, lib.backend, fused))


# ---------------------------------------------------------------- fused multi-head attention
def _attention_numpy(q, k, v, mask, heads, scale):
    """float64 statement of Transformer::MultiHead's core (split heads, scaled scores + additive
    mask, softmax over keys, weighted values, join heads) and its gradient via the closed form."""
    B, Tq, d = q.shape
    Tk = k.shape[1]
    dk = d // heads
    qh = q.reshape(B, Tq, heads, dk).transpose(0, 2, 1, 3).astype(np.float64)
    kh = k.reshape(B, Tk, heads, dk).transpose(0, 2, 1, 3).astype(np.float64)
    vh = v.reshape(B, Tk, heads, dk).transpose(0, 2, 1, 3).astype(np.float64)
    s = scale * qh @ kh.transpose(0, 1, 3, 2)
    if mask is not None:
        s = s + mask.reshape(B, 1, -1, Tk)
    s = s - s.max(-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(-1, keepdims=True)
    o = (p @ vh).transpose(0, 2, 1, 3).reshape(B, Tq, d)
    return o, p, (qh, kh, vh)


@pytest.mark.parametrize("B,H,Tq,Tk,dk,mask_kind", [
    (64, 8, 50, 50, 64, "key"),       # config B encoder self-attention
    (64, 8, 50, 50, 64, "causal"),    # config B decoder self-attention
    (5, 4, 13, 29, 16, "key"),        # cross attention, ragged
    (3, 2, 33, 7, 32, "none"),
    (2, 16, 80, 80, 64, "causal"),    # config E head shape
    (2, 3, 128, 128, 64, "key"),      # largest supported sequence
    # warp-private dk = 64 kernels (swizzled tiles, register-resident strips): ragged edges
    (4, 4, 37, 64, 64, "key"),        # Tq != Tk, full 64-key tile
    (3, 2, 33, 51, 64, "none"),       # odd Tk: scalar probs / mask accesses
    (2, 2, 100, 40, 64, "key"),       # two 64-row query blocks forward, generic kernel backward
    (2, 8, 64, 64, 64, "causal"),     # exactly one full tile, per-query mask
    (3, 1, 5, 7, 64, "key"),          # a single partly filled strip
    (2, 2, 56, 49, 64, "key"),        # 56-row tiles (four CTAs per SM), odd Tk
])
@pytest.mark.parametrize("exact,tol", [(1, 3e-5), (0, 3e-3)], ids=["3xtf32", "tf32"])
def test_multi_head_attention(cuda, oracle, B, H, Tq, Tk, dk, mask_kind, exact, tol):
    d = H * dk
    q, k, v = rnd(1, B, Tq, d), rnd(2, B, Tk, d), rnd(3, B, Tk, d)
    adj = rnd(4, B, Tq, d)
    scale = 1.0 / np.sqrt(dk)
    rs = np.random.RandomState(5)
    if mask_kind == "key":
        keep = (rs.rand(B, 1, 1, Tk) > 0.3).astype(np.float32)
        keep[..., 0] = 1
    elif mask_kind == "causal":
        assert Tq == Tk
        keep = np.tril(np.ones((Tq, Tk), dtype=np.float32))[None, None] * (rs.rand(B, 1, 1, Tk) > 0.2)
        keep[..., 0] = 1
        keep = np.ascontiguousarray(np.broadcast_to(keep, (B, 1, Tq, Tk))).astype(np.float32)
    else:
        keep = None
    mask = None if keep is None else ((1 - keep) * -99999999.0).astype(np.float32)

    def fn(lib):
        out, probs = lib.zeros((B, Tq, d)), lib.zeros((B, H, Tq, Tk))
        qa, ka, va = lib.array(q), lib.array(k), lib.array(v)
        mt = lib.array(mask).t() if mask is not None else None
        lib.call("mrn_multi_head_attention", out.t(), probs.t(), qa.t(), ka.t(), va.t(), mt, H, scale, exact)
        dq, dk_, dv = lib.array(rnd(6, B, Tq, d)), lib.array(rnd(7, B, Tk, d)), lib.array(rnd(8, B, Tk, d))
        lib.call("mrn_multi_head_attention_grad", dq.t(), dk_.t(), dv.t(), lib.array(adj).t(), out.t(), probs.t(), qa.t(), ka.t(), va.t(), H, scale, exact)
        return {"out": out.numpy(), "probs": probs.numpy(), "dq": dq.numpy(), "dk": dk_.numpy(), "dv": dv.numpy()}

    got, exp = both(cuda, oracle, fn)
    for key in exp:
        close(got[key], exp[key], tol, key)
    # independent float64 statement
    o, p, (qh, kh, vh) = _attention_numpy(q, k, v, mask, H, scale)
    close(got["out"], o, tol, "out vs numpy")
    close(got["probs"], p, tol, "probs vs numpy")
    do = adj.reshape(B, Tq, H, dk).transpose(0, 2, 1, 3).astype(np.float64)
    dv = p.transpose(0, 1, 3, 2) @ do
    dp = do @ vh.transpose(0, 1, 3, 2)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True))
    dq = scale * ds @ kh
    dk_ = scale * ds.transpose(0, 1, 3, 2) @ qh
    join = lambda x, T: x.transpose(0, 2, 1, 3).reshape(B, T, d)  # noqa: E731
    close(got["dq"], join(dq, Tq) + rnd(6, B, Tq, d), tol, "dq vs numpy")
    close(got["dk"], join(dk_, Tk) + rnd(7, B, Tk, d), tol, "dk vs numpy")
    close(got["dv"], join(dv, Tk) + rnd(8, B, Tk, d), tol, "dv vs numpy")


@pytest.mark.parametrize("B,H,T,dk", [(3, 4, 50, 64), (2, 2, 80, 64), (2, 2, 21, 32)])
def test_multi_head_attention_aliased_gradients(cuda, oracle, B, H, T, dk):
    # keys and values from the SAME tensor (attention without projections): the gradient tensor is
    # shared too, the kernels must accumulate both contributions into it
    d = H * dk
    q, kv, adj = rnd(11, B, T, d), rnd(12, B, T, d), rnd(13, B, T, d)
    scale = 1.0 / np.sqrt(dk)

    def fn(lib):
        out, probs = lib.zeros((B, T, d)), lib.zeros((B, H, T, T))
        qa, ka = lib.array(q), lib.array(kv)
        lib.call("mrn_multi_head_attention", out.t(), probs.t(), qa.t(), ka.t(), ka.t(), None, H, scale, 1)
        dq, dkv = lib.zeros((B, T, d)), lib.zeros((B, T, d))
        lib.call("mrn_multi_head_attention_grad", dq.t(), dkv.t(), dkv.t(), lib.array(adj).t(), out.t(), probs.t(), qa.t(), ka.t(), ka.t(), H, scale, 1)
        return {"out": out.numpy(), "dq": dq.numpy(), "dkv": dkv.numpy()}

    got, exp = both(cuda, oracle, fn)
    for key in exp:
        close(got[key], exp[key], 3e-5, key)
    o, p, (qh, kh, vh) = _attention_numpy(q, kv, kv, None, H, scale)
    do = adj.reshape(B, T, H, dk).transpose(0, 2, 1, 3).astype(np.float64)
    dp = do @ vh.transpose(0, 1, 3, 2)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True))
    dkv = p.transpose(0, 1, 3, 2) @ do + scale * ds.transpose(0, 1, 3, 2) @ qh
    close(got["dkv"], dkv.transpose(0, 2, 1, 3).reshape(B, T, d), 3e-5, "dK + dV vs numpy")


def test_layer_norm_reference_golden_input(cuda, oracle, goldens):
    # same numbers as the reference's unit test, fed directly to the kernel
    import ctypes  # noqa: F401

    exp = np.array(goldens["operator/layer_norm"]["expected"], dtype=np.float32)
    got = cuda.golden("operator/layer_norm")
    close(got, exp, 2e-5, "ln golden")


# ---------------------------------------------------------------- GRU / LSTM / highway
@pytest.mark.parametrize("rows,cols,with_mask,final", [(64, 128, True, False), (5, 33, False, True), (8, 1024, True, True), (64, 1024, False, False), (3, 132, True, False),
                                                        (70, 1024, True, True), (3200, 256, True, False)])
def test_gru(cuda, oracle, rows, cols, with_mask, final):
    def fn(lib):
        state, xW, sU = lib.array(rnd(1, rows, cols)), lib.array(rnd(2, rows, 3 * cols)), lib.array(rnd(3, rows, 3 * cols))
        b = lib.array(rnd(4, 1, 3 * cols))
        ins = [state, xW, sU, b]
        if with_mask:
            ins.append(lib.array((np.random.RandomState(5).rand(rows, 1) > 0.3).astype(np.float32)))
        out = lib.zeros((rows, cols))
        tl = lib.tensor_list([x.t() for x in ins])
        lib.call("mrn_gru_fast_forward", out.t(), tl, len(ins), int(final))
        adj = lib.array(rnd(6, rows, cols))
        gs = [lib.array(rnd(7, rows, cols)), lib.array(rnd(8, rows, 3 * cols)), lib.array(rnd(9, rows, 3 * cols)), lib.array(rnd(10, 1, 3 * cols))]
        lib.call("mrn_gru_fast_backward", lib.tensor_list([g.t() for g in gs]), tl, len(ins), adj.t(), int(final))
        return {"out": out.numpy(), "gstate": gs[0].numpy(), "gxW": gs[1].numpy(), "gsU": gs[2].numpy(), "gb": gs[3].numpy()}

    compare(cuda, oracle, fn, rtol=5e-5)


@pytest.mark.parametrize("rows,cols,with_mask", [(64, 128, True), (5, 33, False)])
def test_lstm(cuda, oracle, rows, cols, with_mask):
    def fn(lib):
        cell, xW, sU = lib.array(rnd(1, rows, cols)), lib.array(rnd(2, rows, 4 * cols)), lib.array(rnd(3, rows, 4 * cols))
        b = lib.array(rnd(4, 1, 4 * cols))
        ins = [cell, xW, sU, b]
        if with_mask:
            ins.append(lib.array((np.random.RandomState(5).rand(rows, 1) > 0.3).astype(np.float32)))
        tl = lib.tensor_list([x.t() for x in ins])
        c2, h = lib.zeros((rows, cols)), lib.zeros((rows, cols))
        lib.call("mrn_lstm_cell_forward", c2.t(), tl, len(ins))
        tl4 = lib.tensor_list([c2.t(), xW.t(), sU.t(), b.t()])
        lib.call("mrn_lstm_output_forward", h.t(), tl4, 4)
        adj = lib.array(rnd(6, rows, cols))
        gs = [lib.array(rnd(7, rows, cols)), lib.array(rnd(8, rows, 4 * cols)), lib.array(rnd(9, rows, 4 * cols)), lib.array(rnd(10, 1, 4 * cols))]
        lib.call("mrn_lstm_cell_backward", lib.tensor_list([g.t() for g in gs]), tl, len(ins), adj.t())
        lib.call("mrn_lstm_output_backward", lib.tensor_list([g.t() for g in gs]), tl4, 4, adj.t())
        return {"c": c2.numpy(), "h": h.numpy(), "g0": gs[0].numpy(), "g1": gs[1].numpy(), "g2": gs[2].numpy(), "g3": gs[3].numpy()}

    compare(cuda, oracle, fn, rtol=5e-5)


def test_highway(cuda, oracle):
    def fn(lib):
        a, b, t, adj = (lib.array(rnd(i, 33, 65)) for i in range(4))
        out = lib.zeros((33, 65))
        lib.call("mrn_highway_forward", out.t(), a.t(), b.t(), t.t())
        o1, o2, o3 = lib.array(rnd(5, 33, 65)), lib.array(rnd(6, 33, 65)), lib.array(rnd(7, 33, 65))
        lib.call("mrn_highway_backward", o1.t(), o2.t(), o3.t(), a.t(), b.t(), t.t(), adj.t())
        return {"out": out.numpy(), "o1": o1.numpy(), "o2": o2.numpy(), "o3": o3.numpy()}

    compare(cuda, oracle, fn)


# ---------------------------------------------------------------- Bahdanau attention
@pytest.mark.parametrize("T,B,K", [(50, 64, 256), (7, 3, 33), (50, 64, 2048), (9, 5, 132)])
def test_att(cuda, oracle, T, B, K):
    def fn(lib):
        va, ctx, state = lib.array(rnd(1, K, 1)), lib.array(rnd(2, T, B, K)), lib.array(rnd(3, 1, 1, B, K))
        out = lib.zeros((1, T, B, 1))
        lib.call("mrn_att", out.t(), va.t(), ctx.t(), state.t())
        adj = lib.array(rnd(4, 1, T, B, 1).reshape(1, T, B, 1))
        gva, gctx, gst = lib.array(rnd(5, K, 1)), lib.array(rnd(6, T, B, K)), lib.array(rnd(7, 1, 1, B, K))
        lib.call("mrn_att_back", gva.t(), gctx.t(), gst.t(), va.t(), ctx.t(), state.t(), adj.t())
        return {"out": out.numpy(), "gva": gva.numpy(), "gctx": gctx.numpy(), "gst": gst.numpy()}

    compare(cuda, oracle, fn, rtol=5e-5)


# ---------------------------------------------------------------- data movement
@pytest.mark.parametrize("shape,axes", [((1, 50, 64, 32), (0, 2, 1, 3)), ((64, 50, 8, 16), (0, 2, 1, 3)), ((33, 77), (1, 0)),
                                        ((2, 1, 2, 2), (1, 3, 2, 0)), ((2, 1, 2, 2), (2, 0, 1, 3)), ((5, 6, 7), (0, 2, 1)), ((3, 5, 7, 9), (3, 2, 1, 0)),
                                        ((4, 3, 6), (1, 0, 2))])
def test_transpose(cuda, oracle, shape, axes):
    oshape = tuple(shape[a] for a in axes)

    def fn(lib):
        import ctypes

        x = lib.array(rnd(1, *shape))
        out = lib.array(rnd(2, *oshape))  # ASSIGNED, previous content must vanish
        ax = (ctypes.c_int * len(axes))(*axes)
        lib.call("mrn_transpose_nd", out.t(), x.t(), ax)
        return {"out": out.numpy()}

    got, exp = both(cuda, oracle, fn)
    assert np.array_equal(got["out"], exp["out"])
    assert np.array_equal(exp["out"], np.transpose(rnd(1, *shape), axes))


@pytest.mark.parametrize("shape,axis,n", [((1, 2, 2, 3), 2, 4), ((1, 2, 2, 3), -1, 4), ((1, 2, 2, 3), -3, 4), ((1, 2, 2, 3), 0, 4), ((1, 64, 128), -3, 50), ((64, 100), -1, 3),
                                          ((1, 8, 64), -3, 100), ((1, 3, 5), -3, 7), ((6, 5), -1, 97), ((1, 64, 1024), -3, 50)])
def test_concatenate_roundtrip(cuda, oracle, shape, axis, n):
    ax = axis if axis >= 0 else len(shape) + axis
    oshape = list(shape)
    oshape[ax] *= n

    def fn(lib):
        ins = [lib.array(rnd(i, *shape)) for i in range(n)]
        out = lib.zeros(oshape)
        lib.call("mrn_concatenate", out.t(), lib.tensor_list([x.t() for x in ins]), n, axis)
        backs = [lib.array(rnd(100 + i, *shape)) for i in range(n)]
        lib.call("mrn_deconcatenate", lib.tensor_list([x.t() for x in backs]), n, out.t(), axis)
        r = {"out": out.numpy()}
        for i, b in enumerate(backs):
            r["back%d" % i] = b.numpy()
        return r

    got, exp = both(cuda, oracle, fn)
    for k in exp:
        assert np.array_equal(got[k], exp[k]), k
    assert np.array_equal(got["out"], np.concatenate([rnd(i, *shape) for i in range(n)], axis=ax))
    for i in range(n):  # split(concat(x)) == x
        assert np.array_equal(got["back%d" % i], rnd(i, *shape))


@pytest.mark.parametrize("vocab,dim,n", [(1000, 512, 3200), (37, 5, 11)])
def test_rows_gather_scatter(cuda, oracle, vocab, dim, n):
    idx = np.random.RandomState(3).randint(0, vocab, size=n).astype(np.int32)
    idx[: n // 4] = idx[0]  # repeated rows: scatter must accumulate

    def fn(lib):
        table = lib.array(rnd(1, vocab, dim))
        di = lib.array(idx, dtype=np.int32)
        out = lib.zeros((n, dim))
        lib.call("mrn_copy_rows", out.t(), table.t(), di.ptr, n)
        grad = lib.array(rnd(2, vocab, dim))
        adj = lib.array(rnd(3, n, dim))
        lib.call("mrn_paste_rows", grad.t(), adj.t(), di.ptr, n)
        return {"out": out.numpy(), "grad": grad.numpy()}

    got, exp = both(cuda, oracle, fn)
    assert np.array_equal(got["out"], exp["out"])
    close(got["grad"], exp["grad"], 1e-5, "paste_rows")


def test_shift(cuda, oracle):
    import ctypes

    def fn(lib):
        x = lib.array(rnd(1, 10, 4, 8))
        out, back = lib.array(rnd(2, 10, 4, 8)), lib.array(rnd(3, 10, 4, 8))
        sh = (ctypes.c_int * 3)(1, 0, 0)
        lib.call("mrn_shift", out.t(), x.t(), sh, 0)
        lib.call("mrn_shift", back.t(), out.t(), sh, 1)
        return {"out": out.numpy(), "back": back.numpy()}

    got, exp = both(cuda, oracle, fn)
    assert np.array_equal(got["out"], exp["out"]) and np.array_equal(got["back"], exp["back"])
    x = rnd(1, 10, 4, 8)
    assert np.array_equal(got["out"][1:], x[:-1]) and not got["out"][0].any()


# ---------------------------------------------------------------- norm + Adam
def test_l2norm_and_adam(cuda, oracle):
    import ctypes

    n = 100_003 * 4

    def fn(lib):
        p, g = lib.array(rnd(1, 1, n)), lib.array(rnd(2, 1, n, scale=0.01))
        m, v = lib.zeros((1, n)), lib.zeros((1, n))
        norm = ctypes.c_float()
        lib.call("mrn_l2norm", g.t(), ctypes.byref(norm))
        for t in (1, 2, 3):
            lib.call("mrn_adam_step", p.t(), g.t(), m.t(), v.t(), 1e-4, 0.9, 0.999, 1e-8, t, 0.5, 1.0)
        lib.synchronize()
        return {"norm": np.array([norm.value]), "p": p.numpy(), "m": m.numpy(), "v": v.numpy()}

    compare(cuda, oracle, fn, rtol=2e-5)


def test_sgd_and_adagrad_steps(cuda, oracle):
    """gSgd / gAdagrad (reference optimizers/optimizers.cu:7-41) with and without norm clipping:
    CUDA vs the oracle and vs a float64 restatement of the reference's update sequence."""
    n = 50_001 * 4
    P, G = rnd(1, 1, n), rnd(2, 1, n, scale=0.01)

    for clip in (0.0, 0.5):
        def fn(lib):
            p, g, gt = lib.array(P), lib.array(G), lib.zeros((1, n))
            p2 = lib.array(P)
            for _ in range(3):
                lib.call("mrn_sgd_step", p.t(), g.t(), 0.05, 0.5, clip)
                lib.call("mrn_adagrad_step", p2.t(), g.t(), gt.t(), 0.05, 1e-8, 0.5, clip)
            lib.synchronize()
            return {"sgd": p.numpy(), "adagrad": p2.numpy(), "gt": gt.numpy()}

        compare(cuda, oracle, fn, rtol=2e-5)
        # float64 restatement: Norm::clip(g * scale) then the plain updates
        g = G.astype(np.float64) * 0.5
        norm = np.sqrt((g * g).sum())
        if clip > 0 and norm >= clip:
            g = g * (clip / norm)
        p, p2, gt = P.astype(np.float64), P.astype(np.float64), np.zeros_like(g)
        for _ in range(3):
            p = p - 0.05 * g
            gt = gt + g * g
            p2 = p2 - 0.05 / (np.sqrt(gt) + 1e-8) * g
        got = fn(cuda)
        close(got["sgd"], p, 2e-5, "sgd vs float64")
        close(got["adagrad"], p2, 2e-5, "adagrad vs float64")
        close(got["gt"], gt, 2e-5, "adagrad accumulator vs float64")


@pytest.mark.parametrize("p", [0.1, 0.3, 0.5])
def test_dropout_statistics(cuda, oracle, p):
    """Dropout mask (reference kernels/dropout.cu:25-42): values are exactly {0, 1/(1-p)}, the keep
    rate is 1-p within 5 sigma, E[mask] = 1, different seeds give different masks, no visible
    correlation between neighbours.  The oracle's mask obeys the same law (the stream itself is unpinned)."""
    n = 1 << 20
    for lib in (cuda, oracle):
        m = lib.zeros((1, n))
        lib.call("mrn_dropout", m.t(), p, 1234)
        lib.synchronize()
        a = m.numpy().ravel()
        scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
        keep = a != 0
        assert np.all(np.isclose(a[keep], scale, rtol=1e-6)), "kept entries must equal 1/(1-p)"
        sigma = np.sqrt(p * (1 - p) / n)
        assert abs(keep.mean() - (1 - p)) < 5 * sigma, (keep.mean(), 1 - p)
        assert abs(a.mean() - 1.0) < 5 * sigma * scale
        # lag-1 and lag-32 autocorrelation of the keep indicator
        k = keep.astype(np.float64) - keep.mean()
        for lag in (1, 32, 1024):
            c = float((k[:-lag] * k[lag:]).mean() / k.var())
            assert abs(c) < 6.0 / np.sqrt(n), (lag, c)
        m2 = lib.zeros((1, n))
        lib.call("mrn_dropout", m2.t(), p, 1235)
        lib.synchronize()
        b = m2.numpy().ravel() != 0
        agree = (b == keep).mean()
        assert abs(agree - (p * p + (1 - p) * (1 - p))) < 6.0 / np.sqrt(n) + 1e-3, agree
