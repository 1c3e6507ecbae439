"""The drop-in boundary, proven by a build: the REFERENCE's own model code - /root/reference/src/models/transformer.h
and src/models/s2s.h, compiled where they lie and unchanged (oracle/Makefile target `refmodels`; their "marian.h"
umbrella is redirected by oracle/ref_shims/marian.h) - runs on this repo's ExpressionGraph / Node / operator / layer /
rnn API and produces the same parameters (names, shapes, creation order), cost, logits and gradients as this repo's own
model classes (csrc/models/transformer.h, s2s.h).  Both libraries sit on the CPU operator layer, so the comparison
isolates the MODEL code: an error in this repo's rewrite of the reference's graph construction shows up here.
"""
import os

import numpy as np
import pytest

from conftest import ROOT

REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libmarian_oracle_refmodels.so")

CONFIGS = {
    "transformer": "type=transformer;dim-vocabs=60,70;dim-emb=32;transformer-heads=4;transformer-dim-ffn=48;enc-depth=2;dec-depth=2",
    "transformer-tied": "type=transformer;dim-vocabs=64,64;dim-emb=32;transformer-heads=2;transformer-dim-ffn=40;enc-depth=1;dec-depth=2;tied-embeddings-all=true",
    "s2s-gru": "type=s2s;dim-vocabs=60,70;dim-emb=24;dim-rnn=40;enc-depth=2;dec-depth=2",
    "s2s-lstm": "type=s2s;dim-vocabs=60,70;dim-emb=24;dim-rnn=40;enc-cell=lstm;dec-cell=lstm",
    "s2s-deep": "type=s2s;dim-vocabs=60,70;dim-emb=24;dim-rnn=40;enc-depth=3;dec-depth=3;enc-cell-depth=2;dec-cell-base-depth=3;dec-cell-high-depth=2;skip=true;layer-normalization=true",
}


@pytest.fixture(scope="module")
def refmodels(pkg):
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libmarian_oracle_refmodels.so not built (make -C oracle refmodels needs /root/reference)")
    return pkg.Library(REF_LIB)


def run(lib, opts, steps=3):
    t = lib.trainer(opts + ";workspace=128;learn-rate=0.001;gemm-mode=0;graph-replay=false")
    out = {"costs": []}
    for s in range(steps):
        t.next_synthetic_batch(6, 7, 9, padded=True)
        t.compute_gradients(keep_logits=(s == 0))
        if s == 0:
            out["names"] = t.param_names()
            out["logits"] = t.get_tensor("logits")
            out["grads"] = {n: t.get_tensor(n, grad=True) for n, _ in out["names"]}
        t.update()
        out["costs"].append(t.cost())
    out["params"] = t.arena_numpy("params")
    t.close()
    return out


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_reference_model_code_matches_this_repos(oracle, refmodels, name):
    ours = run(oracle, CONFIGS[name])
    ref = run(refmodels, CONFIGS[name])
    # the reference's code creates the same parameters in the same order (-> same flat arena, same init stream)
    assert ref["names"] == ours["names"]
    assert np.allclose(ref["costs"], ours["costs"], rtol=1e-6, atol=0), (ref["costs"], ours["costs"])
    scale = float(np.abs(ref["logits"]).max())
    assert float(np.abs(ref["logits"] - ours["logits"]).max()) <= 1e-6 * scale
    gscale = max(float(np.abs(g).max()) for g in ref["grads"].values())
    for n, g in ref["grads"].items():
        assert float(np.abs(g - ours["grads"][n]).max()) <= 1e-6 * max(gscale, 1e-12), n
    # three Adam updates: weights with analytically zero gradients (attention key biases) follow rounding noise by
    # +-lr per step (the two model codes order a few additions differently), all others agree
    diff = np.abs(ref["params"] - ours["params"])
    assert diff.max() <= 2 * 3 * 1e-3 + 1e-5 and np.mean(diff > 2e-5) < 0.01 and np.median(diff) < 1e-7, (diff.max(), np.mean(diff > 2e-5))


@pytest.mark.parametrize("name", ["transformer", "transformer-tied", "s2s-gru", "s2s-deep"])
def test_reference_model_code_decodes_like_this_repos(oracle, refmodels, name):
    """Beam search (csrc/translator/beam_search.h) over the REFERENCE's decoder code - its TransformerState::select,
    its single-step DecoderTransformer::step / DecoderS2S::step with cached states - against this repo's model classes:
    same n-best lists.  Parameters are trained with this repo's model code and handed over through a checkpoint."""
    import tempfile

    opts = CONFIGS[name] + ";workspace=128;learn-rate=0.003;gemm-mode=0;graph-replay=false"
    t = oracle.trainer(opts)
    for _ in range(200):
        t.next_synthetic_batch(6, 5, 6, padded=True)
        t.compute_gradients()
        t.update()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "model.npz")
        t.save(path)
        r = refmodels.trainer(opts)
        r.load(path)
    t.next_synthetic_batch(5, 6, 6, padded=True)
    src, src_mask = t.get_batch(0)
    trg, trg_mask = t.get_batch(1)
    r.set_batch(src, src_mask, trg, trg_mask)
    for o in ("beam-size=4;normalize=0.6", "beam-size=1;allow-unk=true", "beam-size=3;beam-fused-nth=false"):
        ours, ref = t.translate(o, n_best=3), r.translate(o, n_best=3)
        assert [[h[0] for h in s] for s in ours] == [[h[0] for h in s] for s in ref], o
        assert np.allclose([h[1] for s in ours for h in s], [h[1] for s in ref for h in s], rtol=1e-5), o
        assert o != "beam-size=4;normalize=0.6" or any(len(h[0]) > 1 for s in ours for h in s)  # more than one decoding step
    t.close()
    r.close()
