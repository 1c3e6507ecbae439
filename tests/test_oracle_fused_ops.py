"""CPU (oracle) pin of the two fused operators this repo adds to the operator API:

  multi_head_attention(q, k, v, mask)    ==  split heads / bdot / + mask / softmax / bdot / join heads
  residual_layer_norm(x, r, gamma, beta) ==  layer_norm(x + r, gamma, beta)

The reference (src/models/transformer.h:58-261) only has the right-hand node sequences.  Here the
ORACLE builds the same Transformer twice - from the reference's unfused nodes and from the fused
operators' CPU restatements - and must produce the same loss, logits and parameter gradients.
(The GPU parity tests then compare the fused CUDA kernels with the unfused oracle graph.)
"""
import numpy as np

OPTS = ("type=transformer;dim-vocabs=120,130;dim-emb=32;transformer-heads=4;transformer-dim-ffn=64;"
        "enc-depth=2;dec-depth=2;workspace=128;gemm-mode=0;graph-replay=false")


def _run(oracle, fused):
    flag = "true" if fused else "false"
    t = oracle.trainer(OPTS + ";transformer-fused-attention=%s;transformer-fused-residual-norm=%s" % (flag, flag))
    t.next_synthetic_batch(6, 9, 11, padded=True)
    t.compute_gradients(keep_logits=True)
    out = {"cost": t.cost(), "logits": t.get_tensor("logits"), "grads": {n: t.get_tensor(n, grad=True) for n, _ in t.param_names()}}
    t.update()
    out["params"] = t.arena_numpy("params")
    t.close()
    return out


def test_fused_operators_equal_reference_node_sequences(oracle):
    a, b = _run(oracle, True), _run(oracle, False)
    assert abs(a["cost"] - b["cost"]) <= 1e-6 * abs(b["cost"])
    assert np.allclose(a["logits"], b["logits"], rtol=0, atol=2e-5 * np.abs(b["logits"]).max())
    gscale = max(float(np.abs(g).max()) for g in b["grads"].values())
    for name, g in b["grads"].items():
        err = float(np.abs(a["grads"][name] - g).max())
        assert err <= 2e-5 * max(float(np.abs(g).max()), 1e-2 * gscale), (name, err)
    assert np.abs(a["params"] - b["params"]).max() <= 2.1e-4  # one Adam step of lr 1e-4: sign flips of ~0 gradients only


def test_grouped_product_statement(oracle):
    """mrn_prod_grouped_nt on the oracle: C = beta C + sum_g A_g B_g^T, the chain of accumulating products the
    K-grouped tensor-core launch replaces (AffineNodeOp backward of projections that share their input)."""
    rs = np.random.RandomState(7)
    M, N, K = 37, 24, 40
    for G in (1, 2, 3):
        for beta in (0.0, 1.0):
            As = [rs.standard_normal((M, K)).astype(np.float32) for _ in range(G)]
            Bs = [rs.standard_normal((N, K)).astype(np.float32) for _ in range(G)]
            C0 = rs.standard_normal((M, N)).astype(np.float32)
            exp = beta * C0.astype(np.float64) + sum(a.astype(np.float64) @ b.astype(np.float64).T for a, b in zip(As, Bs))
            g = oracle.gemm(0)
            c = oracle.array(C0)
            a = [oracle.array(x) for x in As]
            b = [oracle.array(x) for x in Bs]
            oracle.call("mrn_prod_grouped_nt", g.h, c.t(), oracle.tensor_list([x.t() for x in a]), oracle.tensor_list([x.t() for x in b]), G, beta)
            assert np.abs(c.numpy() - exp).max() <= 1e-5 * np.abs(exp).max(), (G, beta)
