"""Independent pinning of the oracle where the reference's unit tests are silent
(backward passes, cross-entropy, Cost, bdot, Affine, GRU/LSTM
backward, highway, Adam, clipping have NO golden vectors in the reference).

Every check here compares the oracle (through the C ABI) with float64 PyTorch
autograd on formulas written down independently from the reference's
documentation of the op - so a transcription error in oracle/ cannot hide.
The last test re-implements the whole Marian Transformer forward pass in
PyTorch from the parameter names alone and compares loss, logits and ALL
parameter gradients of a training step."""
import ctypes
import math

import numpy as np
import pytest
import torch

torch.set_default_dtype(torch.float64)


def rnd(seed, *shape, scale=1.0):
    return (np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32)


def close(a, b, tol, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(1e-6, np.abs(b).max())
    err = np.abs(a - b).max() / scale
    assert err <= tol, "%s: %.3e" % (what, err)


def T(x, grad=True):
    return torch.tensor(np.asarray(x, np.float64), requires_grad=grad)


# ------------------------------------------------------------------ row ops
def test_softmax_and_grads(oracle):
    lib = oracle
    x, adj = rnd(1, 7, 13, scale=2), rnd(2, 7, 13)
    X = T(x)
    for name, fwd, bwd, tfn in (("softmax", "mrn_softmax", "mrn_softmax_grad", lambda t: torch.softmax(t, -1)),
                                ("logsoftmax", "mrn_logsoftmax", "mrn_logsoftmax_grad", lambda t: torch.log_softmax(t, -1))):
        out = lib.zeros((7, 13))
        if name == "softmax":
            lib.call(fwd, out.t(), lib.array(x).t(), None)
        else:
            lib.call(fwd, out.t(), lib.array(x).t())
        Y = tfn(X)
        close(out.numpy(), Y.detach().numpy(), 2e-6, name)
        g = lib.zeros((7, 13))
        lib.call(bwd, g.t(), lib.array(adj).t(), out.t())
        (G,) = torch.autograd.grad(Y, X, T(adj, False))
        close(g.numpy(), G.numpy(), 5e-6, name + " grad")


def test_masked_softmax(oracle):
    lib = oracle
    x = rnd(1, 3, 2, 5, 5, scale=2)
    m = (np.random.RandomState(3).rand(3, 1, 1, 5) > 0.4).astype(np.float32)
    m[..., 0] = 1
    out = lib.zeros(x.shape)
    lib.call("mrn_softmax", out.t(), lib.array(x).t(), lib.array(m).t())
    X = torch.tensor(x.astype(np.float64)).masked_fill(torch.tensor(np.broadcast_to(m, x.shape).copy()) == 0, -math.inf)
    close(out.numpy(), torch.softmax(X, -1).numpy(), 2e-6, "masked softmax")


def test_cross_entropy_and_grad(oracle):
    lib = oracle
    rows, cols = 9, 37
    x, adj = rnd(1, rows, cols, scale=2), rnd(2, rows, 1)
    pick = np.random.RandomState(3).randint(0, cols, size=(rows, 1))
    out = lib.zeros((rows, 1))
    lib.call("mrn_cross_entropy_pick", out.t(), lib.array(x).t(), lib.array(pick.astype(np.float32)).t())
    X = T(x)
    CE = torch.nn.functional.cross_entropy(X, torch.tensor(pick[:, 0]), reduction="none")
    close(out.numpy()[:, 0], CE.detach().numpy(), 2e-6, "ce")
    g = lib.zeros((rows, cols))
    lib.call("mrn_cross_entropy_pick_backward", g.t(), lib.array(adj).t(), lib.array(x).t(), lib.array(pick.astype(np.float32)).t())
    (G,) = torch.autograd.grad(CE, X, T(adj[:, 0], False))
    close(g.numpy(), G.numpy(), 5e-6, "ce grad")


@pytest.mark.parametrize("with_beta,eps", [(True, 1e-6), (False, 1e-9)])
@pytest.mark.parametrize("uniform_gamma", [True, False])
def test_layer_norm_and_grad(oracle, with_beta, eps, uniform_gamma):
    """Forward vs autograd always.  Backward: the reference's kernel
    (tensor_operators.cu:1543-1643) leaves gamma OUT of the two row sums and
    multiplies by gamma[id] at the end, which equals the true gradient only when
    gamma is constant along the row (true at initialisation: gamma = 1).  Parity
    means reproducing THAT formula: for non-uniform gamma the oracle is checked
    against the reference formula evaluated in float64, for uniform gamma also
    against autograd."""
    lib = oracle
    rows, cols = 6, 24
    x, beta, adj = rnd(1, rows, cols), 0.3 * rnd(3, 1, cols), rnd(4, rows, cols)
    gamma = np.full((1, cols), 1.3, np.float32) if uniform_gamma else (1 + 0.2 * rnd(2, 1, cols)).astype(np.float32)
    y = lib.zeros((rows, cols))
    bt = lib.array(beta) if with_beta else None
    lib.call("mrn_layer_norm", y.t(), lib.array(x).t(), lib.array(gamma).t(), bt.t() if bt else None, eps)
    X, G, B = T(x), T(gamma), T(beta)
    mu = X.mean(-1, keepdim=True)
    var = ((X - mu) ** 2).mean(-1, keepdim=True)  # biased variance, eps inside the root
    Y = G * (X - mu) / torch.sqrt(var + eps) + (B if with_beta else 0)
    close(y.numpy(), Y.detach().numpy(), 2e-6, "ln")
    gx, gg, gb = lib.zeros((rows, cols)), lib.zeros((1, cols)), lib.zeros((1, cols))
    lib.call("mrn_layer_norm_grad", gx.t(), gg.t(), gb.t() if with_beta else None, lib.array(adj).t(), y.t(), lib.array(x).t(),
             lib.array(gamma).t(), bt.t() if bt else None, eps)
    grads = torch.autograd.grad(Y, [X, G] + ([B] if with_beta else []), T(adj, False))
    # gamma / beta gradients are exact in the reference
    close(gg.numpy(), grads[1].numpy(), 2e-5, "ln dgamma")
    if with_beta:
        close(gb.numpy(), grads[2].numpy(), 2e-5, "ln dbeta")
    # dx: the reference formula in float64
    x64, a64, g64 = x.astype(np.float64), adj.astype(np.float64), gamma.astype(np.float64)
    mu64 = x64.mean(-1, keepdims=True)
    sigma = np.sqrt(eps + ((x64 - mu64) ** 2).mean(-1, keepdims=True))
    xhat = (x64 - mu64) / sigma
    ref_dx = g64 * (cols * a64 - a64.sum(-1, keepdims=True) - (a64 * xhat).sum(-1, keepdims=True) * xhat) / (cols * sigma)
    close(gx.numpy(), ref_dx, 2e-5, "ln dx (reference formula)")
    if uniform_gamma:
        close(gx.numpy(), grads[0].numpy(), 2e-5, "ln dx (autograd)")


# ------------------------------------------------------------------ cells
@pytest.mark.parametrize("final", [False, True])
def test_gru_cell_and_grads(oracle, final):
    """GRU gate math as documented in SURVEY.md 2.3 #20."""
    lib = oracle
    rows, cols = 5, 8
    s, xW, sU, b = rnd(1, rows, cols), rnd(2, rows, 3 * cols), rnd(3, rows, 3 * cols), rnd(4, 1, 3 * cols)
    m = (np.random.RandomState(5).rand(rows, 1) > 0.4).astype(np.float32)
    adj = rnd(6, rows, cols)
    ins = [lib.array(v) for v in (s, xW, sU, b, m)]
    tl = lib.tensor_list([v.t() for v in ins])
    out = lib.zeros((rows, cols))
    lib.call("mrn_gru_fast_forward", out.t(), tl, 5, int(final))
    S, XW, SU, B, M = T(s), T(xW), T(sU), T(b), T(m, False)
    C = cols
    r = torch.sigmoid(XW[:, :C] + SU[:, :C] + B[:, :C])
    z = torch.sigmoid(XW[:, C:2 * C] + SU[:, C:2 * C] + B[:, C:2 * C])
    if final:
        h = torch.tanh(XW[:, 2 * C:] + (SU[:, 2 * C:] + B[:, 2 * C:]) * r)
    else:
        h = torch.tanh(XW[:, 2 * C:] + SU[:, 2 * C:] * r + B[:, 2 * C:])
    O = M * ((1 - z) * h + z * S) + (1 - M) * S
    close(out.numpy(), O.detach().numpy(), 2e-6, "gru fwd")
    gs = [lib.zeros((rows, cols)), lib.zeros((rows, 3 * cols)), lib.zeros((rows, 3 * cols)), lib.zeros((1, 3 * cols))]
    lib.call("mrn_gru_fast_backward", lib.tensor_list([g.t() for g in gs]), tl, 5, lib.array(adj).t(), int(final))
    G = torch.autograd.grad(O, [S, XW, SU, B], T(adj, False))
    for got, exp, name in zip(gs, G, ("dstate", "dxW", "dsU", "db")):
        close(got.numpy(), exp.numpy(), 1e-5, "gru " + name)


def test_lstm_cell_and_grads(oracle):
    lib = oracle
    rows, cols = 5, 8
    c, xW, sU, b = rnd(1, rows, cols), rnd(2, rows, 4 * cols), rnd(3, rows, 4 * cols), rnd(4, 1, 4 * cols)
    m = (np.random.RandomState(5).rand(rows, 1) > 0.4).astype(np.float32)
    adjc, adjh = rnd(6, rows, cols), rnd(7, rows, cols)
    ins = [lib.array(v) for v in (c, xW, sU, b, m)]
    tl = lib.tensor_list([v.t() for v in ins])
    c2, h = lib.zeros((rows, cols)), lib.zeros((rows, cols))
    lib.call("mrn_lstm_cell_forward", c2.t(), tl, 5)
    tl4 = lib.tensor_list([c2.t(), ins[1].t(), ins[2].t(), ins[3].t()])
    lib.call("mrn_lstm_output_forward", h.t(), tl4, 4)
    Cc, XW, SU, B, M = T(c), T(xW), T(sU), T(b), T(m, False)
    C = cols
    pre = XW + SU + B
    gf, gi, gc, go = torch.sigmoid(pre[:, :C]), torch.sigmoid(pre[:, C:2 * C]), torch.tanh(pre[:, 2 * C:3 * C]), torch.sigmoid(pre[:, 3 * C:])
    C2 = M * (gf * Cc + gi * gc) + (1 - M) * Cc
    H = go * torch.tanh(C2)
    close(c2.numpy(), C2.detach().numpy(), 2e-6, "lstm c")
    close(h.numpy(), H.detach().numpy(), 2e-6, "lstm h")
    # total gradient of sum(adjh*H) + sum(adjc*C2) wrt inputs = output-kernel grads (incl. d/dC2) chained into the cell kernel
    gC2 = lib.array(adjc)  # d loss / d C2 from elsewhere
    gs_o = [gC2, lib.zeros((rows, 4 * cols)), lib.zeros((rows, 4 * cols)), lib.zeros((1, 4 * cols))]
    lib.call("mrn_lstm_output_backward", lib.tensor_list([g.t() for g in gs_o]), tl4, 4, lib.array(adjh).t())
    gs_c = [lib.zeros((rows, cols)), gs_o[1], gs_o[2], gs_o[3]]
    lib.call("mrn_lstm_cell_backward", lib.tensor_list([g.t() for g in gs_c]), tl, 5, gC2.t())
    G = torch.autograd.grad((H * T(adjh, False)).sum() + (C2 * T(adjc, False)).sum(), [Cc, XW, SU, B])
    for got, exp, name in zip(gs_c, G, ("dcell", "dxW", "dsU", "db")):
        close(got.numpy(), exp.numpy(), 1e-5, "lstm " + name)


def test_bahdanau_attention_and_grads(oracle):
    lib = oracle
    Tn, Bn, K = 6, 3, 10
    va, ctx, st, adj = rnd(1, K, 1), rnd(2, Tn, Bn, K), rnd(3, 1, 1, Bn, K), rnd(4, 1, Tn, Bn, 1)
    out = lib.zeros((1, Tn, Bn, 1))
    lib.call("mrn_att", out.t(), lib.array(va).t(), lib.array(ctx).t(), lib.array(st).t())
    VA, CTX, ST = T(va), T(ctx), T(st)
    O = (torch.tanh(CTX + ST.reshape(1, Bn, K)) * VA.reshape(1, 1, K)).sum(-1)
    close(out.numpy().reshape(Tn, Bn), O.detach().numpy(), 2e-6, "att")
    gva, gctx, gst = lib.zeros((K, 1)), lib.zeros((Tn, Bn, K)), lib.zeros((1, 1, Bn, K))
    lib.call("mrn_att_back", gva.t(), gctx.t(), gst.t(), lib.array(va).t(), lib.array(ctx).t(), lib.array(st).t(), lib.array(adj).t())
    G = torch.autograd.grad(O, [VA, CTX, ST], T(adj.reshape(Tn, Bn), False))
    close(gva.numpy(), G[0].numpy(), 1e-5, "att dva")
    close(gctx.numpy(), G[1].numpy(), 1e-5, "att dctx")
    close(gst.numpy(), G[2].numpy(), 1e-5, "att dstate")


def test_elementwise_grads(oracle):
    lib = oracle
    x, adj = rnd(1, 4, 9), rnd(2, 4, 9)
    X = T(x)
    for fwd, bwd, tfn, nin in (("swish", "swish_grad", lambda t: t * torch.sigmoid(t), 3), ("logit", "logit_grad", torch.sigmoid, 2),
                               ("relu", "relu_grad", torch.relu, 2)):
        out = lib.zeros(x.shape)
        lib.call("mrn_element", fwd.encode(), out.t(), lib.tensor_list([lib.array(x).t()]), 1, 0.0)
        Y = tfn(X)
        close(out.numpy(), Y.detach().numpy(), 2e-6, fwd)
        g = lib.zeros(x.shape)
        ins = {"swish_grad": [lib.array(adj), lib.array(x), out], "logit_grad": [lib.array(adj), out], "relu_grad": [lib.array(adj), lib.array(x)]}[bwd]
        lib.call("mrn_add", bwd.encode(), 1.0, g.t(), lib.tensor_list([i.t() for i in ins]), len(ins), 0.0)
        (G,) = torch.autograd.grad(Y, X, T(adj, False))
        close(g.numpy(), G.numpy(), 5e-6, bwd)


def test_adam_with_clipping(oracle):
    """clip-norm + Adam as specified in SURVEY.md 8a (a20)."""
    lib = oracle
    n = 1000
    p0, g = rnd(1, 1, n), rnd(2, 1, n, scale=0.2)
    p, m, v = lib.array(p0), lib.zeros((1, n)), lib.zeros((1, n))
    P = p0.astype(np.float64).copy()
    M, V = np.zeros_like(P), np.zeros_like(P)
    eta, b1, b2, eps, clip, scale = 1e-2, 0.9, 0.999, 1e-8, 1.0, 0.5
    for t in (1, 2, 3):
        lib.call("mrn_adam_step", p.t(), lib.array(g).t(), m.t(), v.t(), eta, b1, b2, eps, t, scale, clip)
        G = g.astype(np.float64) * scale
        norm = np.sqrt((G ** 2).sum())
        if norm >= clip:
            G = G * clip / norm
        M = b1 * M + (1 - b1) * G
        V = b2 * V + (1 - b2) * G * G
        P = P - eta * (M / (1 - b1 ** t)) / (np.sqrt(V / (1 - b2 ** t)) + eps)
    close(p.numpy(), P, 2e-6, "adam params")
    close(m.numpy(), M, 1e-5, "adam m")


# ------------------------------------------------------------------ whole model
def marian_transformer_torch(params, src, smask, trg, tmask, heads, depth):
    """Marian's Transformer (models/transformer.h) re-stated in PyTorch from parameter names."""
    P = params
    Ts, B = src.shape
    Tt = trg.shape[0]
    d = P["encoder_Wemb"].shape[1]

    def pos(T):
        nts = d // 2
        inc = math.log(10000.0) / (nts - 1.0)
        v = torch.arange(T).reshape(-1, 1) * torch.exp(torch.arange(nts) * -inc).reshape(1, -1)
        return torch.cat([torch.sin(v), torch.cos(v)], -1)  # [T, d]

    def ln(x, pre):
        mu = x.mean(-1, keepdim=True)
        var = ((x - mu) ** 2).mean(-1, keepdim=True)
        return P[pre + "_ln_scale"] * (x - mu) / torch.sqrt(var + 1e-6) + P[pre + "_ln_bias"]

    def mha(pre, q, kv, addmask):
        def proj(x, n):
            y = x @ P[pre + "_W" + n] + P[pre + "_b" + n]
            return y.reshape(y.shape[0], y.shape[1], heads, d // heads).transpose(1, 2)  # [B,H,T,dk]

        qh, kh, vh = proj(q, "q"), proj(kv, "k"), proj(kv, "v")
        w = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(d // heads) + addmask, -1)
        o = (w @ vh).transpose(1, 2).reshape(q.shape[0], q.shape[1], d)
        return o @ P[pre + "_Wo"] + P[pre + "_bo"]

    def block_att(pre, x, kv, addmask):
        return ln(mha(pre, x, kv, addmask) + x, pre + "_Wo")

    def block_ffn(pre, x):
        h = x @ P[pre + "_W1"] + P[pre + "_b1"]
        h = h * torch.sigmoid(h)
        h = h @ P[pre + "_W2"] + P[pre + "_b2"]
        return ln(h + x, pre + "_ffn")

    # encoder: [B, Ts, d]
    x = (math.sqrt(d) * P["encoder_Wemb"][src] + pos(Ts).unsqueeze(1)).transpose(0, 1)
    enc_add = ((1 - smask.transpose(0, 1)) * -99999999.0).reshape(B, 1, 1, Ts)
    for i in range(1, depth + 1):
        x = block_att("encoder_l%d_self" % i, x, x, enc_add)
        x = block_ffn("encoder_l%d_ffn" % i, x)
    ctx = x
    # decoder: teacher forcing with the target embeddings shifted by one step
    y = P["decoder_Wemb"][trg]  # [Tt, B, d]
    y = torch.cat([torch.zeros_like(y[:1]), y[:-1]], 0)
    q = (math.sqrt(d) * y + pos(Tt).unsqueeze(1)).transpose(0, 1)
    tri = torch.tril(torch.ones(Tt, Tt))
    self_add = ((1 - tri.reshape(1, Tt, Tt) * tmask.transpose(0, 1).reshape(B, 1, Tt)) * -99999999.0).reshape(B, 1, Tt, Tt)
    for i in range(1, depth + 1):
        q = block_att("decoder_l%d_self" % i, q, q, self_add)
        q = block_att("decoder_l%d_context" % i, q, ctx, enc_add)
        q = block_ffn("decoder_l%d_ffn" % i, q)
    logits = q.transpose(0, 1) @ P["decoder_ff_logit_out_W"] + P["decoder_ff_logit_out_b"]  # [Tt, B, V]
    ce = torch.nn.functional.cross_entropy(logits.reshape(Tt * B, -1), trg.reshape(-1), reduction="none").reshape(Tt, B)
    cost = (ce * tmask).sum(0).mean()  # ce-mean: sum over time, mean over sentences
    return cost, logits


def test_transformer_step_against_independent_torch_model(oracle):
    heads, depth, d = 4, 2, 32
    opts = ("type=transformer;dim-vocabs=60,70;dim-emb=%d;transformer-heads=%d;transformer-dim-ffn=48;enc-depth=%d;dec-depth=%d;"
            "workspace=64;clip-norm=0" % (d, heads, depth, depth))
    rs = np.random.RandomState(0)
    B, Ts, Tt = 5, 7, 6
    src, trg = rs.randint(2, 60, size=(Ts, B)), rs.randint(2, 70, size=(Tt, B))
    src[-1] = 0
    trg[-1] = 0
    sm, tm = np.ones((Ts, B), np.float32), np.ones((Tt, B), np.float32)
    sm[-2:, 1] = 0
    tm[-3:, 2] = 0

    t = oracle.trainer(opts)
    t.set_batch(src, sm, trg, tm)
    t.compute_gradients(keep_logits=True)
    cost = t.cost()
    names = t.param_names()
    params = {n: torch.tensor(t.get_tensor(n).reshape(s).astype(np.float64), requires_grad=True) for n, s in names}
    C, L = marian_transformer_torch(params, torch.tensor(src), torch.tensor(sm.astype(np.float64)), torch.tensor(trg), torch.tensor(tm.astype(np.float64)), heads, depth)
    assert abs(cost - C.item()) <= 2e-6 * abs(C.item()), (cost, C.item())
    close(t.get_tensor("logits"), L.detach().numpy().reshape(-1), 5e-6, "logits")
    grads = torch.autograd.grad(C, [params[n] for n, _ in names])
    # key biases have an analytically ZERO gradient (softmax is invariant to a shift of all
    # scores of a row): compare every tensor on the scale of the largest gradients, not its own
    gscale = max(float(G.abs().max()) for G in grads)
    for (n, _), G in zip(names, grads):
        got = t.get_tensor(n, grad=True).astype(np.float64)
        exp = G.numpy().reshape(-1)
        scale = max(np.abs(exp).max(), 1e-2 * gscale)
        assert np.abs(got - exp).max() <= 5e-5 * scale, (n, np.abs(got - exp).max(), scale)
    t.close()


# ------------------------------------------------------------------ whole s2s (RNN + attention) step
def marian_s2s_torch(params, src, smask, trg, tmask, enc_depth, dec_depth, cell_type="gru"):
    """Marian's deep RNN encoder-decoder with Bahdanau attention (src/models/s2s.h, src/rnn/{rnn,cells,attention}.h,
    enc-type bidirectional) re-stated in PyTorch from the parameter names and the papers' formulas alone:
      GRU:            r, z = sigmoid(x W_{r,z} + s U_{r,z} + b_{r,z});  h~ = tanh(x W_x + (s U_x) r + b_x)
                      ("final" cells of the conditional GRU: h~ = tanh(x W_x + (s U_x + b_x) r));  s' = (1 - z) h~ + z s
      LSTM:           f, i, o = sigmoid(.), g = tanh(.) of x W + h U + b (column blocks f | i | g | o);  c' = f c + i g;
                      h' = o tanh(c');  a padded step keeps c (h is recomputed from the kept c with this step's o)
      encoder:        two full-depth stacks, one reading left-to-right, one right-to-left (padded steps keep the state),
                      top outputs concatenated
      decoder start:  tanh(W mean_t(context) + b), the same vector for every layer (output and cell state)
      conditional:    s1 = CELL1(y_{t-1}, s);  c_t = sum_j softmax_j(v . tanh(W_c ctx_j + W_s s1 + b)) ctx_j;  s' = CELL2(c_t, s1)
      readout:        logits = W2 tanh(W0 y + b0 + W1 s_top + b1 + W2 c + b2) + b
    A state is the pair (output, cell); GRUs carry no cell."""
    P = params
    Ts, B = src.shape
    Tt = trg.shape[0]
    lstm = cell_type == "lstm"

    def cell(name, x_proj, state, mask=None, final=False):
        s, c = state
        D = s.shape[-1]
        if lstm:
            g = s @ P[name + "_U"] + P[name + "_b"]
            if x_proj is not None:
                g = g + x_proj
            c2 = torch.sigmoid(g[..., :D]) * c + torch.sigmoid(g[..., D:2 * D]) * torch.tanh(g[..., 2 * D:3 * D])
            if mask is not None:
                c2 = mask * c2 + (1 - mask) * c
            return torch.sigmoid(g[..., 3 * D:]) * torch.tanh(c2), c2
        U = torch.cat([P[name + "_U"], P[name + "_Ux"]], -1)
        b = torch.cat([P[name + "_b"], P[name + "_bx"]], -1)
        su = s @ U
        xp = x_proj if x_proj is not None else torch.zeros_like(su)
        r = torch.sigmoid(xp[..., :D] + su[..., :D] + b[..., :D])
        z = torch.sigmoid(xp[..., D:2 * D] + su[..., D:2 * D] + b[..., D:2 * D])
        if final:
            h = torch.tanh(xp[..., 2 * D:] + (su[..., 2 * D:] + b[..., 2 * D:]) * r)
        else:
            h = torch.tanh(xp[..., 2 * D:] + su[..., 2 * D:] * r + b[..., 2 * D:])
        out = (1 - z) * h + z * s
        return (out if mask is None else mask * out + (1 - mask) * s), None

    def in_proj(name, x):
        return x @ (P[name + "_W"] if lstm else torch.cat([P[name + "_W"], P[name + "_Wx"]], -1))

    def layer(name, xs, start, mask, backward=False):
        xp = in_proj(name, xs)  # all time steps at once
        T = xs.shape[0]
        state = start
        outs = [None] * T
        for t in (range(T - 1, -1, -1) if backward else range(T)):
            state = cell(name, xp[t], state, None if mask is None else mask[t])
            outs[t] = state[0]
        return torch.stack(outs, 0)

    D = P["encoder_bi_U"].shape[0]
    x = P["encoder_Wemb"][src]  # [Ts, B, E]
    m = smask.unsqueeze(-1)     # [Ts, B, 1]
    zeros = (torch.zeros(B, D), torch.zeros(B, D))

    def stack(base, backward):
        h = x
        for l in range(1, enc_depth + 1):
            name = base if l == 1 else "%s_l%d_cell1" % (base, l)
            h = layer(name, h, zeros, m, backward)
        return h

    ctx = torch.cat([stack("encoder_bi", False), stack("encoder_bi_r", True)], -1)  # [Ts, B, 2D]

    mean_ctx = (ctx * m).sum(0) / m.sum(0)
    start = torch.tanh(mean_ctx @ P["decoder_ff_state_W"] + P["decoder_ff_state_b"])

    y = P["decoder_Wemb"][trg]
    y = torch.cat([torch.zeros_like(y[:1]), y[:-1]], 0)  # teacher forcing: embeddings shifted by one step

    mapped_ctx = ctx @ P["decoder_Wc_att"] + P["decoder_b_att"]  # [Ts, B, 2D]
    xp1 = in_proj("decoder_cell1", y)
    state = (start, start)
    tops, contexts = [], []
    for t in range(Tt):
        s1 = cell("decoder_cell1", xp1[t], state)
        score = (torch.tanh(mapped_ctx + (s1[0] @ P["decoder_W_comb_att"]).unsqueeze(0)) @ P["decoder_U_att"]).squeeze(-1)  # [Ts, B]
        e = torch.softmax(score.masked_fill(smask == 0, -math.inf), 0)
        c = (e.unsqueeze(-1) * ctx).sum(0)  # [B, 2D]
        state = cell("decoder_cell2", in_proj("decoder_cell2", c), s1, final=True)
        tops.append(state[0])
        contexts.append(c)
    h = torch.stack(tops, 0)
    for l in range(2, dec_depth + 1):
        h = layer("decoder_l%d_cell1" % l, h, (start, start), None)
    c_all = torch.stack(contexts, 0)
    hid = torch.tanh(y @ P["decoder_ff_logit_l1_W0"] + P["decoder_ff_logit_l1_b0"] + h @ P["decoder_ff_logit_l1_W1"] + P["decoder_ff_logit_l1_b1"]
                     + c_all @ P["decoder_ff_logit_l1_W2"] + P["decoder_ff_logit_l1_b2"])
    logits = hid @ P["decoder_ff_logit_l2_W"] + P["decoder_ff_logit_l2_b"]
    ce = torch.nn.functional.cross_entropy(logits.reshape(Tt * B, -1), trg.reshape(-1), reduction="none").reshape(Tt, B)
    return (ce * tmask).sum(0).mean(), logits


@pytest.mark.parametrize("cell_type,enc_depth,dec_depth", [("gru", 1, 1), ("gru", 2, 3), ("lstm", 1, 1), ("lstm", 2, 2)])
def test_s2s_step_against_independent_torch_model(oracle, cell_type, enc_depth, dec_depth):
    """The RNN + attention half of the hot path: loss, logits and ALL parameter gradients of one training step of the
    deep GRU s2s model (the architecture of BASELINE.json configs[0] / configs[2]) against float64 PyTorch autograd on
    an implementation written from the formulas - pins the graph-level backward of GRU (incl. the "final" cell) and LSTM cells,
    masked bidirectional stacks, the attention cell input, the multi-input readout and the parameter naming."""
    opts = ("type=s2s;dim-vocabs=60,70;dim-emb=16;dim-rnn=24;enc-depth=%d;dec-depth=%d;enc-cell=%s;dec-cell=%s;workspace=64;clip-norm=0"
            % (enc_depth, dec_depth, cell_type, cell_type))
    rs = np.random.RandomState(1)
    B, Ts, Tt = 4, 6, 5
    src, trg = rs.randint(2, 60, size=(Ts, B)), rs.randint(2, 70, size=(Tt, B))
    src[-1] = 0
    trg[-1] = 0
    sm, tm = np.ones((Ts, B), np.float32), np.ones((Tt, B), np.float32)
    sm[-2:, 1] = 0   # padded source positions (mask path of both directions and of the attention softmax)
    sm[-3:, 3] = 0
    tm[-2:, 2] = 0

    t = oracle.trainer(opts)
    t.set_batch(src, sm, trg, tm)
    t.compute_gradients(keep_logits=True)
    cost = t.cost()
    names = t.param_names()
    params = {n: torch.tensor(t.get_tensor(n).reshape(s).astype(np.float64), requires_grad=True) for n, s in names}
    C, L = marian_s2s_torch(params, torch.tensor(src), torch.tensor(sm.astype(np.float64)), torch.tensor(trg), torch.tensor(tm.astype(np.float64)), enc_depth, dec_depth, cell_type)
    assert abs(cost - C.item()) <= 2e-6 * abs(C.item()), (cost, C.item())
    close(t.get_tensor("logits"), L.detach().numpy().reshape(-1), 5e-6, "logits")
    grads = torch.autograd.grad(C, [params[n] for n, _ in names], allow_unused=True)
    gscale = max(float(G.abs().max()) for G in grads if G is not None)
    for (n, _), G in zip(names, grads):
        assert G is not None, "the torch model does not use parameter " + n
        got = t.get_tensor(n, grad=True).astype(np.float64)
        exp = G.numpy().reshape(-1)
        scale = max(np.abs(exp).max(), 1e-2 * gscale)
        assert np.abs(got - exp).max() <= 5e-5 * scale, (n, np.abs(got - exp).max(), scale)
    t.close()
