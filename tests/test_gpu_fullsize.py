"""Parity at the sizes BASELINE.json names, CUDA vs the CPU oracle on the IDENTICAL synthetic
batch (V = 32000, data seed 1111) and the identical initialisation (seed 1234):

  * Transformer-base 64 x 50 (configs[1]), dense and padded: step-1 cost, a logits slab, the costs
    after three clip+Adam updates - within the north-star 1e-4 in the exact modes (0 = fp32 SIMT,
    2 = bf16x3 tcgen05); the throughput modes (3 = tf32, 4 = bf16 operands) are bounded at their
    own, stated, tolerance and their error is printed;
  * deep GRU s2s 4+4 64 x 50 (configs[2]): same checks;
  * Transformer-big geometry (d = 1024, 16 heads, ffn 4096, T = 80: configs[4]) on 16 sentences.

The oracle runs each configuration once (module-scoped fixtures, a few seconds per step on the
GPU box's host cores).  Reference code paths: src/models/transformer.h:384-449,495-662,
src/layers/generic.cpp:5-42, src/models/s2s.h, src/rnn/rnn.h:54-143.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROW_STRIDE = 4  # logits rows compared: every 4th (800 x 32000 for 64 x 50)
V = 32000

GRU_OPTS = {"type": "s2s", "dim-vocabs": [V, V], "dim-emb": 512, "dim-rnn": 1024, "enc-depth": 4, "dec-depth": 4,
            "enc-cell": "gru", "dec-cell": "gru", "cost-type": "ce-mean", "label-smoothing": 0, "optimizer": "adam",
            "learn-rate": 0.0001, "clip-norm": 1, "seed": 1234, "workspace": 16384}


def big_opts(pkg):
    o = pkg.transformer_base_options(gemm_mode=0, workspace=16384)
    o.update({"dim-emb": 1024, "transformer-heads": 16, "transformer-dim-ffn": 4096})
    return o


def run(lib, opts, mode, batch, steps, padded=False, replay=False):
    o = dict(opts)
    o["gemm-mode"] = mode
    o["graph-replay"] = "true" if replay else "false"
    o["data-seed"] = 1111
    t = lib.trainer(o)
    out = {"costs": []}
    for s in range(steps):
        t.next_synthetic_batch(batch[0], batch[1], batch[2], padded=padded)
        t.compute_gradients(keep_logits=(s == 0))
        if s == 0:
            out["cost0"] = t.cost()
            out["logits"] = t.get_tensor("logits").reshape(-1, V)[::ROW_STRIDE].copy()
        t.update()
        out["costs"].append(t.cost())
    t.close()
    return out


def check(got, exp, tol, what, steps_tol=None):
    cost_err = abs(got["cost0"] - exp["cost0"]) / abs(exp["cost0"])
    scale = max(1e-6, float(np.abs(exp["logits"]).max()))
    logit_err = float(np.abs(got["logits"].astype(np.float64) - exp["logits"]).max()) / scale
    n = min(len(got["costs"]), len(exp["costs"]))
    step_err = float(np.max(np.abs(np.array(got["costs"][:n]) - np.array(exp["costs"][:n])) / np.abs(exp["costs"][:n])))
    print("\n[parity %s] cost rel err %.3e, logits max rel err %.3e, costs over %d updates rel err %.3e (tolerance %.1e)"
          % (what, cost_err, logit_err, n, step_err, tol))
    assert cost_err <= tol, (what, got["cost0"], exp["cost0"])
    assert logit_err <= tol, (what, logit_err)
    assert step_err <= (steps_tol or 3 * tol), (what, got["costs"], exp["costs"])


# ---------------------------------------------------------------- Transformer-base 64 x 50
@pytest.fixture(scope="module")
def oracle_tb(oracle, pkg):
    return run(oracle, pkg.transformer_base_options(gemm_mode=0), 0, (64, 50, 50), 3)


@pytest.mark.parametrize("mode,tol", [(0, 1e-4), (2, 1e-4), (3, 3e-3), (4, 2e-2)], ids=["fp32", "bf16x3", "tf32", "bf16"])
def test_transformer_base_full_size_matches_oracle(cuda, pkg, oracle_tb, mode, tol):
    got = run(cuda, pkg.transformer_base_options(gemm_mode=mode), mode, (64, 50, 50), 3)
    check(got, oracle_tb, tol, "transformer-base 64x50 mode %d" % mode)


def test_transformer_base_full_size_replay_matches_oracle(cuda, pkg, oracle_tb):
    """The same three updates through CUDA-graph capture + replay (what bench.py times), exact mode."""
    o = pkg.transformer_base_options(gemm_mode=2)
    o["graph-replay"] = "true"
    o["data-seed"] = 1111
    t = cuda.trainer(o)
    costs = []
    for s in range(3):
        t.next_synthetic_batch(64, 50, 50)
        t.compute_gradients()
        t.update()
        costs.append(t.cost())
    st = t.stats()
    t.close()
    assert st["plans"] == 1 and st["replays"] >= 1, st
    assert np.allclose(costs, oracle_tb["costs"], rtol=3e-4), (costs, oracle_tb["costs"])


@pytest.fixture(scope="module")
def oracle_tb_padded(oracle, pkg):
    return run(oracle, pkg.transformer_base_options(gemm_mode=0), 0, (64, 50, 50), 1, padded=True)


@pytest.mark.parametrize("mode,tol", [(2, 1e-4), (3, 3e-3), (4, 2e-2)], ids=["bf16x3", "tf32", "bf16"])
def test_transformer_base_padded_batch_matches_oracle(cuda, pkg, oracle_tb_padded, mode, tol):
    """Lengths uniform in [T/2, T], padded and target-length sorted: the mask path at full size."""
    got = run(cuda, pkg.transformer_base_options(gemm_mode=mode), mode, (64, 50, 50), 1, padded=True)
    check(got, oracle_tb_padded, tol, "transformer-base padded 64x50 mode %d" % mode)


# ---------------------------------------------------------------- deep GRU s2s 64 x 50
@pytest.fixture(scope="module")
def oracle_gru(oracle):
    return run(oracle, GRU_OPTS, 0, (64, 50, 50), 3)


@pytest.mark.parametrize("mode,tol", [(0, 1e-4), (2, 1e-4), (3, 3e-3), (4, 2e-2)], ids=["fp32", "bf16x3", "tf32", "bf16"])
def test_deep_gru_full_size_matches_oracle(cuda, oracle_gru, mode, tol):
    got = run(cuda, GRU_OPTS, mode, (64, 50, 50), 2)
    check(got, oracle_gru, tol, "deep GRU s2s 64x50 mode %d" % mode)


def test_deep_gru_full_size_replay_matches_oracle(cuda, oracle_gru):
    """The same updates through CUDA-graph capture + replay (what bench.py --model s2s-deep-gru times): the captured
    graph carries the two encoder lanes and the side stream as parallel branches.  Exact mode."""
    o = dict(GRU_OPTS)
    o.update({"gemm-mode": 2, "graph-replay": "true", "data-seed": 1111})
    t = cuda.trainer(o)
    costs = []
    for s in range(3):
        t.next_synthetic_batch(64, 50, 50)
        t.compute_gradients()
        t.update()
        costs.append(t.cost())
    st = t.stats()
    t.close()
    assert st["plans"] == 1 and st["replays"] >= 1, st
    assert np.allclose(costs, oracle_gru["costs"], rtol=3e-4), (costs, oracle_gru["costs"])


# ---------------------------------------------------------------- Transformer-big geometry, T = 80
@pytest.fixture(scope="module")
def oracle_big(oracle, pkg):
    return run(oracle, big_opts(pkg), 0, (16, 80, 80), 1)


@pytest.mark.parametrize("mode,tol", [(2, 1e-4), (3, 3e-3), (4, 2e-2)], ids=["bf16x3", "tf32", "bf16"])
def test_transformer_big_geometry_matches_oracle(cuda, pkg, oracle_big, mode, tol):
    """d = 1024, 16 heads (dk = 64), ffn 4096, sequences of 80 tokens: the attention kernels
    for 64 < T <= 128 and the d = 1024 row kernels against the oracle."""
    got = run(cuda, big_opts(pkg), mode, (16, 80, 80), 1)
    check(got, oracle_big, tol, "transformer-big 16x80 mode %d" % mode)
