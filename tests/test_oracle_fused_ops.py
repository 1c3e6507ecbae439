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
