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
