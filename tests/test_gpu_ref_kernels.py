"""Kernel-vs-kernel parity with the REFERENCE'S OWN CUDA kernels.

oracle/_ref/libmarian_ref_kernels.so is /root/reference/src/kernels/tensor_operators.cu
(+ tensors/*.cu) compiled unmodified for sm_100a behind oracle/ref_harness.cu
(`make -C oracle ref`, done by __graft_entry__.build() in the build container;
the GPU box only sees the prebuilt .so).  Same device buffers are handed to
both libraries.  Tolerance 5e-5 relative: the reference is built with
--use_fast_math (__expf, approximate division) and sums in a different order.
"""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT
from test_gpu_ops import close, rnd

pytestmark = pytest.mark.gpu

REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libmarian_ref_kernels.so")
TOL = 5e-5


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return ctypes.CDLL(REF_LIB)


def rc(x):
    assert x == 0, "reference kernel call failed with %d" % x


def test_row_ops_match_reference_kernels(cuda, ref):
    rows, cols = 320, 512
    x = cuda.array(rnd(1, rows, cols, scale=2))
    for mine, theirs in (("mrn_softmax", ref.ref_softmax), ("mrn_logsoftmax", ref.ref_logsoftmax)):
        a, b = cuda.zeros((rows, cols)), cuda.zeros((rows, cols))
        if mine == "mrn_softmax":
            cuda.call(mine, a.t(), x.t(), None)
            rc(theirs(b.t(), x.t(), None))
        else:
            cuda.call(mine, a.t(), x.t())
            rc(theirs(b.t(), x.t()))
        cuda.synchronize()
        close(a.numpy(), b.numpy(), TOL, mine)

    # masked softmax with the attention broadcast pattern [B,1,1,T] on [B,H,T,T]
    shape, mshape = (8, 4, 50, 50), (8, 1, 1, 50)
    m = (np.random.RandomState(3).rand(*mshape) > 0.3).astype(np.float32)
    m[..., 0] = 1
    xs, mk = cuda.array(rnd(2, *shape, scale=2)), cuda.array(m)
    a, b = cuda.zeros(shape), cuda.zeros(shape)
    mt = mk.t()
    cuda.call("mrn_softmax", a.t(), xs.t(), mt)
    rc(ref.ref_softmax(b.t(), xs.t(), ctypes.byref(mt)))
    cuda.synchronize()
    close(a.numpy(), b.numpy(), TOL, "masked softmax")

    # layer norm forward + backward (transformer shape, eps 1e-6)
    gamma, beta, adj = cuda.array(1 + 0.1 * rnd(3, 1, cols)), cuda.array(0.1 * rnd(4, 1, cols)), cuda.array(rnd(5, rows, cols))
    ya, yb = cuda.zeros((rows, cols)), cuda.zeros((rows, cols))
    bt = beta.t()
    cuda.call("mrn_layer_norm", ya.t(), x.t(), gamma.t(), bt, 1e-6)
    rc(ref.ref_layer_norm(yb.t(), x.t(), gamma.t(), ctypes.byref(bt), ctypes.c_float(1e-6)))
    cuda.synchronize()
    close(ya.numpy(), yb.numpy(), TOL, "layer norm")
    outs = []
    for lib_is_ref in (False, True):
        gx, gg, gb = cuda.array(rnd(6, rows, cols)), cuda.array(rnd(7, 1, cols)), cuda.array(rnd(8, 1, cols))
        gbt = gb.t()
        if lib_is_ref:
            rc(ref.ref_layer_norm_grad(gx.t(), gg.t(), ctypes.byref(gbt), adj.t(), yb.t(), x.t(), gamma.t(), ctypes.byref(bt), ctypes.c_float(1e-6)))
        else:
            cuda.call("mrn_layer_norm_grad", gx.t(), gg.t(), gbt, adj.t(), yb.t(), x.t(), gamma.t(), bt, 1e-6)
        cuda.synchronize()
        outs.append((gx.numpy(), gg.numpy(), gb.numpy()))
    for g_mine, g_ref, name in zip(outs[0], outs[1], ("dx", "dgamma", "dbeta")):
        close(g_mine, g_ref, 2e-4, "layer norm " + name)  # atomics in a different order on both sides


def test_cross_entropy_matches_reference_kernels(cuda, ref):
    rows, cols = 64, 32000
    x = cuda.array(rnd(1, rows, cols, scale=2))
    pick = cuda.array(np.random.RandomState(2).randint(0, cols, size=(rows, 1)).astype(np.float32))
    adj = cuda.array(rnd(3, rows, 1))
    a, b = cuda.zeros((rows, 1)), cuda.zeros((rows, 1))
    cuda.call("mrn_cross_entropy_pick", a.t(), x.t(), pick.t())
    rc(ref.ref_cross_entropy_pick(b.t(), x.t(), pick.t()))
    cuda.synchronize()
    close(a.numpy(), b.numpy(), TOL, "ce")
    ga, gb = cuda.zeros((rows, cols)), cuda.zeros((rows, cols))
    cuda.call("mrn_cross_entropy_pick_backward", ga.t(), adj.t(), x.t(), pick.t())
    rc(ref.ref_cross_entropy_pick_backward(gb.t(), adj.t(), x.t(), pick.t()))
    cuda.synchronize()
    close(ga.numpy(), gb.numpy(), TOL, "ce backward")


def test_elementwise_and_reductions_match_reference_kernels(cuda, ref):
    full, small = (1, 50, 64, 512), (1, 512)
    a, b = cuda.array(rnd(1, *full)), cuda.array(rnd(2, *small))
    for f in (b"plus", b"mult"):
        o1, o2 = cuda.zeros(full), cuda.zeros(full)
        tl = cuda.tensor_list([a.t(), b.t()])
        cuda.call("mrn_element", f, o1.t(), tl, 2, 0.0)
        rc(ref.ref_element(f, o2.t(), tl, 2, ctypes.c_float(0.0)))
        cuda.synchronize()
        assert np.array_equal(o1.numpy(), o2.numpy()), f
    # bias gradient (the reference's one-thread-per-column gAddGeneric) and last-axis reduction
    for oshape in ((1, 512), (1, 50, 64, 1)):
        o1, o2 = cuda.array(rnd(3, *oshape)), cuda.array(rnd(3, *oshape))
        tl = cuda.tensor_list([a.t()])
        cuda.call("mrn_add", b"id", 1.0, o1.t(), tl, 1, 0.0)
        rc(ref.ref_add(b"id", ctypes.c_float(1.0), o2.t(), tl, 1, ctypes.c_float(0.0)))
        cuda.synchronize()
        close(o1.numpy(), o2.numpy(), 2e-4, "add " + str(oshape))
    x = cuda.array(rnd(4, 3200, 2048))
    o1, o2 = cuda.zeros((3200, 2048)), cuda.zeros((3200, 2048))
    tl = cuda.tensor_list([x.t()])
    cuda.call("mrn_element", b"swish", o1.t(), tl, 1, 0.0)
    rc(ref.ref_element(b"swish", o2.t(), tl, 1, ctypes.c_float(0.0)))
    cuda.synchronize()
    close(o1.numpy(), o2.numpy(), TOL, "swish")


def test_gru_and_attention_match_reference_kernels(cuda, ref):
    rows, cols = 64, 1024
    ins = [cuda.array(rnd(1, rows, cols)), cuda.array(rnd(2, rows, 3 * cols)), cuda.array(rnd(3, rows, 3 * cols)), cuda.array(rnd(4, 1, 3 * cols)),
           cuda.array((np.random.RandomState(5).rand(rows, 1) > 0.3).astype(np.float32))]
    tl = cuda.tensor_list([v.t() for v in ins])
    adj = cuda.array(rnd(6, rows, cols))
    res = []
    for use_ref in (False, True):
        out = cuda.zeros((rows, cols))
        gs = [cuda.array(rnd(7, rows, cols)), cuda.array(rnd(8, rows, 3 * cols)), cuda.array(rnd(9, rows, 3 * cols)), cuda.array(rnd(10, 1, 3 * cols))]
        gl = cuda.tensor_list([g.t() for g in gs])
        if use_ref:
            rc(ref.ref_gru_fast_forward(out.t(), tl, 5, 0))
            rc(ref.ref_gru_fast_backward(gl, tl, 5, adj.t(), 0))
        else:
            cuda.call("mrn_gru_fast_forward", out.t(), tl, 5, 0)
            cuda.call("mrn_gru_fast_backward", gl, tl, 5, adj.t(), 0)
        cuda.synchronize()
        res.append([out.numpy()] + [g.numpy() for g in gs])
    for mine, theirs, name in zip(res[0], res[1], ("out", "dstate", "dxW", "dsU", "db")):
        close(mine, theirs, 1e-4, "gru " + name)

    T, B, K = 50, 64, 512
    va, ctx, st = cuda.array(rnd(1, K, 1)), cuda.array(rnd(2, T, B, K)), cuda.array(rnd(3, 1, 1, B, K))
    adj = cuda.array(rnd(4, 1, T, B, 1))
    res = []
    for use_ref in (False, True):
        out = cuda.zeros((1, T, B, 1))
        gva, gctx, gst = cuda.array(rnd(5, K, 1)), cuda.array(rnd(6, T, B, K)), cuda.array(rnd(7, 1, 1, B, K))
        if use_ref:
            rc(ref.ref_att(out.t(), va.t(), ctx.t(), st.t()))
            rc(ref.ref_att_back(gva.t(), gctx.t(), gst.t(), va.t(), ctx.t(), st.t(), adj.t()))
        else:
            cuda.call("mrn_att", out.t(), va.t(), ctx.t(), st.t())
            cuda.call("mrn_att_back", gva.t(), gctx.t(), gst.t(), va.t(), ctx.t(), st.t(), adj.t())
        cuda.synchronize()
        res.append([out.numpy(), gva.numpy(), gctx.numpy(), gst.numpy()])
    for mine, theirs, name in zip(res[0], res[1], ("att", "dva", "dctx", "dstate")):
        close(mine, theirs, 2e-4, "att " + name)


def test_gemm_matches_reference_cublas(cuda, ref):
    """The reference's Prod is cublasSgemm (fp32): fp32-SIMT and bf16x3 tensor-core modes must agree with it."""
    M, K, N = 3200, 512, 2048
    A, B = cuda.array(rnd(1, M, K)), cuda.array(rnd(2, K, N))
    cref = cuda.zeros((M, N))
    rc(ref.ref_prod(cref.t(), A.t(), B.t(), 0, 0, ctypes.c_float(0.0), ctypes.c_float(1.0)))
    for mode, tol in ((0, 1e-5), (2, 5e-5), (1, 1.5e-2)):
        g = cuda.gemm(mode)
        c = cuda.zeros((M, N))
        cuda.call("mrn_prod", g.h, c.t(), A.t(), B.t(), 0, 0, 0.0, 1.0)
        cuda.synchronize()
        close(c.numpy(), cref.numpy(), tol, "prod mode %d" % mode)


def test_adam_matches_reference_update_sequence(cuda, ref):
    """Fused clip+Adam kernel vs the reference's L2Norm + scale + three Element passes."""
    n = 1 << 20
    p0, g0 = rnd(1, 1, n), rnd(2, 1, n, scale=0.01)
    pa, ga, ma, va = cuda.array(p0), cuda.array(g0), cuda.zeros((1, n)), cuda.zeros((1, n))
    pb, gb, mb, vb = cuda.array(p0), cuda.array(g0), cuda.zeros((1, n)), cuda.zeros((1, n))
    for t in (1, 2, 3):
        cuda.call("mrn_adam_step", pa.t(), ga.t(), ma.t(), va.t(), 1e-4, 0.9, 0.999, 1e-8, t, 1.0, 1.0)
        gb.upload(g0)  # the reference clips IN PLACE; restore the raw gradient
        rc(ref.ref_adam_step(pb.t(), gb.t(), mb.t(), vb.t(), ctypes.c_float(1e-4), ctypes.c_float(0.9), ctypes.c_float(0.999), ctypes.c_float(1e-8), t, ctypes.c_float(1.0)))
    cuda.synchronize()
    close(pa.numpy(), pb.numpy(), 1e-6, "adam params")
    close(ma.numpy(), mb.numpy(), 1e-5, "adam m")
    close(va.numpy(), vb.numpy(), 1e-5, "adam v")
