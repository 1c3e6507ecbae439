"""Prod / ProdBatched / ProdAffine: the tcgen05 paths (tf32 on the raw fp32 tensors, packed
bf16 and hi/lo-split bf16x3) and the fp32 SIMT path against the CPU oracle and float64 numpy.

Tolerances (relative to the output magnitude):
  fp32 SIMT  1e-5      bf16x3  5e-5 (operands carry 16 mantissa bits)      bf16  1e-2
  tf32  2e-3 (operands carry 11 mantissa bits)
Mode 4 (bf16 SHADOW operands read by TMA in all four transpose cases) has the arithmetic of mode 1
(packed bf16): it is checked against float64 at the bf16 tolerance AND against mode 1 at 2e-5.
"""
import numpy as np
import pytest

from test_gpu_ops import close, rnd

pytestmark = pytest.mark.gpu

TOL = {0: 1e-5, 2: 5e-5, 1: 1.5e-2, 3: 2e-3, 4: 1.5e-2}


def ref_prod(A, B, tA, tB, beta, alpha, C0):
    a = A.reshape(-1, A.shape[-1]).astype(np.float64)
    b = B.reshape(-1, B.shape[-1]).astype(np.float64)
    a = a.T if tA else a
    b = b.T if tB else b
    return alpha * (a @ b) + beta * C0.astype(np.float64)


SHAPES = [  # (A shape, B shape, transA, transB)
    ((2, 2, 3), (3, 2), False, False),           # the reference's own dot test
    ((3200, 512), (512, 512), False, False),     # projection forward (config B)
    ((3200, 512), (512, 512), False, True),      # dX = D W^T
    ((3200, 512), (3200, 512), True, False),     # dW = X^T D   (split-K path)
    ((3200, 2048), (512, 2048), False, True),
    ((300, 70), (70, 130), False, False),        # ragged everything
    ((129, 65), (33, 65), False, True),
    ((65, 129), (65, 33), True, False),
    ((64, 1024), (1024, 3072), False, False),    # RNN step shape
    ((77, 40), (90, 77), True, True),
    ((1, 512), (512, 1000), False, False),
    # 16-byte aligned ragged shapes: stay on the TMA-direct tf32 kernel in mode 3, all four layouts
    ((300, 72), (72, 132), False, False),
    ((132, 68), (36, 68), False, True),
    ((68, 132), (68, 36), True, False),
    ((76, 40), (92, 76), True, True),
    ((3200, 512), (3200, 2048), True, False),    # FFN dW (split-K)
    ((640, 4096), (4096, 512), False, False),    # long reduction
]


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("sa,sb,tA,tB", SHAPES)
@pytest.mark.parametrize("beta,alpha", [(0.0, 1.0), (1.0, 0.125)])
def test_prod(cuda, oracle, mode, sa, sb, tA, tB, beta, alpha):
    A, B = rnd(1, *sa), rnd(2, *sb)
    m = (np.prod(sa) // sa[-1]) if not tA else sa[-1]
    n = sb[-1] if not tB else (np.prod(sb) // sb[-1])
    C0 = rnd(3, int(m), int(n))
    exp = ref_prod(A, B, tA, tB, beta, alpha, C0)

    def run(lib, md):
        g = lib.gemm(md)
        a, b, c = lib.array(A), lib.array(B), lib.array(C0)
        lib.call("mrn_prod", g.h, c.t(), a.t(), b.t(), int(tA), int(tB), beta, alpha)
        lib.synchronize()
        return c.numpy()

    got = run(cuda, mode)
    close(got, exp, TOL[mode], "cuda vs float64")
    if mode == 0 and A.size < 2_000_000:
        close(run(oracle, 0), exp, 1e-5, "oracle vs float64")


@pytest.mark.parametrize("mode", [3, 4])
@pytest.mark.parametrize("M,N,K", [(3200, 2048, 512), (304, 72, 96), (136, 72, 40), (64, 2048, 64)])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_prod_swish_grad_nt(cuda, mode, M, N, K, beta):
    # dH = beta dH + (dY W^T) o swish'(H) in the epilogue of the tf32 product (feed-forward backward pass)
    A, B, H, C0 = rnd(1, M, K), rnd(2, N, K), 2.5 * rnd(3, M, N), rnd(4, M, N)
    h = H.astype(np.float64)
    sg = 1.0 / (1.0 + np.exp(-h))
    exp = beta * C0 + (A.astype(np.float64) @ B.astype(np.float64).T) * (sg * (1.0 + h * (1.0 - sg)))
    g = cuda.gemm(mode)
    c = cuda.array(C0)
    cuda.call("mrn_prod_swish_grad_nt", g.h, c.t(), cuda.array(A).t(), cuda.array(B).t(), cuda.array(H).t(), beta)
    cuda.synchronize()
    close(c.numpy(), exp, TOL[mode], "gated product vs float64")
    # and against the two-step form on the same GPU path (product, then the element-wise swish backward)
    t = cuda.zeros((M, N))
    cuda.call("mrn_prod", g.h, t.t(), cuda.array(A).t(), cuda.array(B).t(), 0, 1, 0.0, 1.0)
    cuda.synchronize()
    two = beta * C0 + t.numpy().astype(np.float64) * (sg * (1.0 + h * (1.0 - sg)))
    close(c.numpy(), two, 2e-5, "gated product vs product + swish'")


@pytest.mark.parametrize("mode", [0, 3, 4])
@pytest.mark.parametrize("M,N,K,G", [
    (3200, 512, 512, 3),   # dX of the q/k/v projections (config B): one K-grouped launch in mode 3
    (3200, 512, 512, 2),   # key/value pair of a cross-attention block
    (300, 72, 96, 3),      # ragged M / N, whole k-blocks
    (129, 64, 40, 2),      # K not a multiple of the k-block: falls back to the chain of products
    (64, 64, 64, 1),
])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_prod_grouped_nt(cuda, oracle, mode, M, N, K, G, beta):
    As = [rnd(10 + g, M, K) for g in range(G)]
    Bs = [rnd(20 + g, N, K) for g in range(G)]
    C0 = rnd(3, M, N)
    exp = beta * C0.astype(np.float64) + sum(a.astype(np.float64) @ b.astype(np.float64).T for a, b in zip(As, Bs))

    def run(lib):
        g = lib.gemm(mode)
        c = lib.array(C0)
        a = [lib.array(x) for x in As]
        b = [lib.array(x) for x in Bs]
        lib.call("mrn_prod_grouped_nt", g.h, c.t(), lib.tensor_list([x.t() for x in a]), lib.tensor_list([x.t() for x in b]), G, beta)
        lib.synchronize()
        return c.numpy()

    close(run(cuda), exp, TOL[mode], "cuda vs float64")
    if mode == 0:
        close(run(oracle), exp, 1e-5, "oracle vs float64")


@pytest.mark.parametrize("mode", [4, 3])
@pytest.mark.parametrize("M,N,K,G,beta", [
    (3200, 512, 2048, 1, 1.0),   # feed-forward dX (config B): 32 k-blocks dealt out over 8 tile columns
    (3200, 512, 512, 3, 1.0),    # q/k/v dX, K-grouped: three bias gradients from one launch
    (3200, 512, 512, 1, 0.0),
    (300, 136, 192, 2, 1.0),     # ragged M / N: more tile columns than k-blocks per group
    (64, 64, 64, 1, 0.0),        # one tile column
    (3200, 1024, 1024, 1, 1.0),  # Transformer-big projection
])
def test_prod_grouped_nt_with_bias_gradients(cuda, mode, M, N, K, G, beta):
    """The input-gradient products of the training step also deliver the bias gradients: column sums of their A
    operands (the adjoints), taken from the A tiles inside the tensor-core kernel - every tile column sums its share
    of the k-blocks.  Sums ACCUMULATE into their targets (gradients)."""
    As = [rnd(10 + g, M, K) for g in range(G)]
    Bs = [rnd(20 + g, N, K) for g in range(G)]
    C0, S0 = rnd(3, M, N), [rnd(30 + g, 1, K) for g in range(G)]
    exp = beta * C0.astype(np.float64) + sum(a.astype(np.float64) @ b.astype(np.float64).T for a, b in zip(As, Bs))
    g = cuda.gemm(mode)
    c = cuda.array(C0)
    a = [cuda.array(x) for x in As]
    b = [cuda.array(x) for x in Bs]
    sums = [cuda.array(x) for x in S0]
    cuda.call("mrn_prod_grouped_nt_sums", g.h, c.t(), cuda.tensor_list([x.t() for x in a]), cuda.tensor_list([x.t() for x in b]), G, beta,
              cuda.tensor_list([x.t() for x in sums]))
    cuda.synchronize()
    close(c.numpy(), exp, TOL[mode], "product vs float64")
    for k in range(G):
        ref = S0[k].astype(np.float64) + As[k].astype(np.float64).sum(axis=0, keepdims=True)
        # the sums are taken from the operand as the tensor core sees it (bf16 in mode 4, fp32 tiles in mode 3)
        # (fp32 accumulation of M values through atomics: 5e-4 of the largest sum)
        close(sums[k].numpy(), ref, 2e-2 if mode == 4 else 5e-4, "bias gradient %d" % k)


@pytest.mark.parametrize("mode", [3, 4])
@pytest.mark.parametrize("M,N,K", [(3200, 2048, 512), (3200, 4096, 1024), (304, 200, 96)])
def test_prod_swish_grad_nt_with_bias_gradient(cuda, mode, M, N, K):
    """Gated product (persistent kernel at full size) + column sums of dY by the column-sum warps: the k-blocks of a
    tile row are dealt out over its tile columns."""
    A, B, H, C0, S0 = rnd(1, M, K), rnd(2, N, K), 2.5 * rnd(3, M, N), rnd(4, M, N), rnd(5, 1, K)
    h = H.astype(np.float64)
    sg = 1.0 / (1.0 + np.exp(-h))
    exp = C0 + (A.astype(np.float64) @ B.astype(np.float64).T) * (sg * (1.0 + h * (1.0 - sg)))
    g = cuda.gemm(mode)
    c, s = cuda.array(C0), cuda.array(S0)
    cuda.call("mrn_prod_swish_grad_nt_sums", g.h, c.t(), cuda.array(A).t(), cuda.array(B).t(), cuda.array(H).t(), 1.0, s.t())
    cuda.synchronize()
    close(c.numpy(), exp, TOL[mode], "gated product vs float64")
    close(s.numpy(), S0.astype(np.float64) + A.astype(np.float64).sum(axis=0, keepdims=True), 2e-2 if mode == 4 else 5e-4, "bias gradient")


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("sa,sb,tA,tB", [
    ((64, 8, 50, 64), (64, 8, 50, 64), False, True),    # Q K^T  (config B)
    ((64, 8, 50, 50), (64, 8, 50, 64), False, False),   # P V
    ((64, 8, 50, 50), (64, 8, 50, 64), True, False),    # backward forms
    ((64, 8, 50, 64), (64, 8, 50, 64), True, False),    # A^T B  [64,64]
    ((3, 2, 7, 5), (3, 2, 5, 9), False, False),
    ((1, 1, 130, 70), (4, 2, 70, 33), False, False),    # A shared by all batches (stride 0)
    ((4, 2, 33, 70), (1, 1, 70, 130), False, False),    # B shared
    ((5, 3, 52, 64), (5, 3, 64, 52), False, False),     # aligned: TMA-direct with batch coordinate
    ((5, 3, 52, 64), (5, 3, 52, 40), True, False),
    ((1, 1, 132, 72), (4, 2, 72, 36), False, False),
])
def test_prod_batched(cuda, oracle, mode, sa, sb, tA, tB):
    A, B = rnd(1, *sa), rnd(2, *sb)
    a64 = A.astype(np.float64).reshape(-1, sa[-2], sa[-1])
    b64 = B.astype(np.float64).reshape(-1, sb[-2], sb[-1])
    if tA:
        a64 = a64.transpose(0, 2, 1)
    if tB:
        b64 = b64.transpose(0, 2, 1)
    C0 = rnd(3, max(a64.shape[0], b64.shape[0]), a64.shape[1], b64.shape[2])
    exp = 0.125 * np.matmul(a64, b64) + C0
    g = cuda.gemm(mode)
    a, b, c = cuda.array(A), cuda.array(B), cuda.array(C0)
    bshape = (max(sa[0], sb[0]), max(sa[1], sb[1]), a64.shape[1], b64.shape[2])
    cuda.call("mrn_prod_batched", g.h, c.t(bshape), a.t(), b.t(), int(tA), int(tB), 1.0, 0.125)
    cuda.synchronize()
    close(c.numpy(), exp, TOL[mode], "batched")


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("M,K,N", [(3200, 512, 2048), (130, 70, 50), (320, 512, 32000)])
def test_prod_affine(cuda, mode, M, K, N):
    A, B, bias = rnd(1, M, K), rnd(2, K, N), rnd(3, 1, N)
    exp = A.astype(np.float64) @ B.astype(np.float64) + bias
    g = cuda.gemm(mode)
    a, b, bb, c = cuda.array(A), cuda.array(B), cuda.array(bias), cuda.array(rnd(4, M, N))
    cuda.call("mrn_prod_affine", g.h, c.t(), a.t(), b.t(), bb.t())
    cuda.synchronize()
    close(c.numpy(), exp, TOL[mode], "affine")


def test_gemm_linearity_full_size(cuda):
    """Size-independent property at the logits shape of config B: A(B1+B2) == AB1 + AB2
    up to bf16x3 rounding, and the tensor-core result agrees with the fp32 SIMT kernel."""
    M, K, N = 3200, 512, 32000
    A, B1, B2 = rnd(1, M, K), rnd(2, K, N, scale=0.05), rnd(3, K, N, scale=0.05)
    outs = {}
    for mode in (0, 2):
        g = cuda.gemm(mode)
        a = cuda.array(A)
        c1, c2, c12 = cuda.zeros((M, N)), cuda.zeros((M, N)), cuda.zeros((M, N))
        cuda.call("mrn_prod", g.h, c1.t(), a.t(), cuda.array(B1).t(), 0, 0, 0.0, 1.0)
        cuda.call("mrn_prod", g.h, c2.t(), a.t(), cuda.array(B2).t(), 0, 0, 0.0, 1.0)
        cuda.call("mrn_prod", g.h, c12.t(), a.t(), cuda.array(B1 + B2).t(), 0, 0, 0.0, 1.0)
        cuda.synchronize()
        outs[mode] = c12.numpy()
        close(c1.numpy() + c2.numpy(), outs[mode], 5e-5, "linearity mode %d" % mode)
    close(outs[2], outs[0], 5e-5, "bf16x3 vs fp32")


BF16_SHAPES = [  # 16-byte aligned bf16 rows (cols % 8 == 0): the TMA-direct bf16 kernel, all four layouts
    ((3200, 512), (512, 512), False, False),
    ((3200, 512), (512, 2048), False, False),
    ((3200, 512), (512, 512), False, True),
    ((3200, 2048), (512, 2048), False, True),
    ((3200, 512), (3200, 512), True, False),     # dW (split-K), both operands MN-major
    ((3200, 512), (3200, 2048), True, False),
    ((304, 72), (72, 136), False, False),         # ragged M / N / K tails through TMA zero fill
    ((136, 72), (40, 72), False, True),
    ((72, 136), (72, 40), True, False),
    ((80, 40), (96, 80), True, True),
    ((64, 1024), (1024, 3072), False, False),    # RNN step shape
]


@pytest.mark.parametrize("sa,sb,tA,tB", BF16_SHAPES)
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_bf16_shadow_mode_equals_packed_bf16(cuda, sa, sb, tA, tB, beta):
    """Mode 4 reads bf16 copies of the row-major operands directly (K-major and MN-major shared-memory
    descriptors); mode 1 packs every operand into K-major bf16 first.  Same rounded operands, same fp32
    accumulation: the results agree to accumulation-order noise."""
    A, B = rnd(1, *sa), rnd(2, *sb)
    m = (np.prod(sa) // sa[-1]) if not tA else sa[-1]
    n = sb[-1] if not tB else (np.prod(sb) // sb[-1])
    C0 = rnd(3, int(m), int(n))
    out = {}
    for mode in (1, 4):
        g = cuda.gemm(mode)
        a, b, c = cuda.array(A), cuda.array(B), cuda.array(C0)
        cuda.call("mrn_prod", g.h, c.t(), a.t(), b.t(), int(tA), int(tB), beta, 0.5)
        cuda.synchronize()
        out[mode] = c.numpy()
    close(out[4], out[1], 2e-5, "bf16 shadows vs packed bf16")


def test_gemm_full_size_logits_shapes_throughput_modes(cuda):
    """The three logits-sized products of config B (forward NN, input gradient NT, weight gradient TN with
    split-K) in the throughput modes 3 (tf32) and 4 (bf16) against the fp32 SIMT kernel."""
    M, K, N = 3200, 512, 32000
    X, W, D = rnd(1, M, K), rnd(2, K, N, scale=0.05), rnd(3, M, N, scale=0.02)
    ref = {}
    for mode in (0, 3, 4):
        g = cuda.gemm(mode)
        x, w, d = cuda.array(X), cuda.array(W), cuda.array(D)
        y, dx, dw = cuda.zeros((M, N)), cuda.zeros((M, K)), cuda.zeros((K, N))
        cuda.call("mrn_prod", g.h, y.t(), x.t(), w.t(), 0, 0, 0.0, 1.0)
        cuda.call("mrn_prod", g.h, dx.t(), d.t(), w.t(), 0, 1, 0.0, 1.0)
        cuda.call("mrn_prod", g.h, dw.t(), x.t(), d.t(), 1, 0, 0.0, 1.0)
        cuda.synchronize()
        got = {"y": y.numpy(), "dx": dx.numpy(), "dw": dw.numpy()}
        if mode == 0:
            ref = got
        else:
            for k in got:
                close(got[k], ref[k], TOL[mode], "mode %d %s" % (mode, k))


@pytest.mark.parametrize("M,K,N,n,tA,beta,bias", [
    (3200, 512, 512, 3, False, 0.0, True),    # q / k / v projections of config B
    (3200, 512, 512, 2, False, 0.0, True),    # key / value of a cross-attention block
    (3200, 512, 512, 3, True, 1.0, False),    # their weight gradients: X^T adj_g, split-K with TMA reduce-add
    (304, 72, 136, 3, False, 0.0, True),      # ragged tiles
    (304, 72, 136, 2, True, 1.0, False),
    (2560, 1024, 1024, 3, False, 0.0, True),  # Transformer-big geometry: 128-wide tiles
])
def test_prod_shared_a(cuda, M, K, N, n, tA, beta, bias):
    """Products that share their A operand as one launch (bf16 shadow mode) vs float64 and vs the single products."""
    import ctypes

    A = rnd(1, M, K)                           # stored [M, K]; the weight-gradient form contracts over its rows (op(A) = A^T)
    Bs = [rnd(10 + i, (M if tA else K), N) for i in range(n)]
    rows = K if tA else M                      # C = op(A) B: [K, N] for the weight-gradient form
    Cs0 = [rnd(20 + i, rows, N) for i in range(n)]
    bs = [rnd(30 + i, 1, N) for i in range(n)] if bias else None
    a64 = A.astype(np.float64)
    exp = [beta * Cs0[i] + (a64.T if tA else a64) @ Bs[i].astype(np.float64) + (bs[i] if bias else 0.0) for i in range(n)]
    out = {}
    for mode in (4, 1):
        g = cuda.gemm(mode)
        cs = [cuda.array(c) for c in Cs0]
        b = [cuda.array(x) for x in Bs]
        bb = [cuda.array(x) for x in bs] if bias else None
        fused = ctypes.c_int(-1)
        cuda.call("mrn_prod_shared_a", g.h, cuda.tensor_list([c.t() for c in cs]), cuda.array(A).t(), cuda.tensor_list([x.t() for x in b]),
                  cuda.tensor_list([x.t() for x in bb]) if bias else None, n, int(tA), beta, ctypes.byref(fused))
        cuda.synchronize()
        out[mode] = [c.numpy() for c in cs]
        assert fused.value == (1 if mode == 4 else 0)
        for i in range(n):
            close(out[mode][i], exp[i], TOL[mode], "shared-A product %d, mode %d" % (i, mode))
    for i in range(n):
        close(out[4][i], out[1][i], 2e-5, "one launch vs single packed products")
