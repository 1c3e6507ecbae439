"""Beam-search decoding on the GPU (csrc/translator/*, csrc/kernels/nth_element.cu) against the CPU oracle.

  * n-best selection kernels: against numpy at the sizes of BASELINE.json's vocabulary (V = 32000, 64 sentences x beam 12),
    ragged / tiny / odd sizes, ties, suppressed word, first step; the generic range variant on ranges longer than one CTA's segment;
  * whole searches: parameters trained on the CPU oracle are loaded into the CUDA trainer through a checkpoint; exact GEMM
    modes return the oracle's n-best lists (words identical, scores to 1e-4), the bf16 mode is checked against its own
    full-sequence training path (the hypothesis score equals minus the cross-entropy of the hypothesis);
  * the fused selection and the reference's node sequence agree on the GPU;
  * Transformer-base dimensions (V = 32000, d = 512, 6+6 layers), 64 sentences, beam 6: runs, and is timed.
"""
import os
import time

import numpy as np
import pytest

from test_translator import S2S, TRANSFORMER, logsoftmax, sequence_logprobs, trained

pytestmark = pytest.mark.gpu


def expected_nbest(logits, prev, batch, per, n, suppress):
    V = logits.shape[-1]
    total = prev.astype(np.float32)[:, None].astype(np.float64) + logsoftmax(logits.astype(np.float64))
    if suppress >= 0:
        total[:, suppress] = np.finfo(np.float32).min
    total = total.reshape(per, batch, V).transpose(1, 0, 2).reshape(batch, per * V)
    keys = np.stack([np.lexsort((np.arange(per * V), -total[s]))[:n] + s * per * V for s in range(batch)])
    costs = np.stack([total[s][keys[s] - s * per * V] for s in range(batch)])
    return costs, keys


@pytest.mark.parametrize("beam,batch,V,n,first,suppress", [
    (12, 64, 32000, 12, False, 1),
    (12, 64, 32000, 12, True, 1),
    (6, 1, 32000, 6, False, -1),
    (5, 3, 37, 5, False, 1),
    (4, 7, 8193, 4, False, -1),
    (3, 2, 50257, 3, False, 1),
    (1, 9, 1000, 1, False, -1),
    (2, 2, 5, 2, False, -1),
])
def test_fused_nbest_kernel_matches_numpy(cuda, beam, batch, V, n, first, suppress):
    rs = np.random.RandomState(beam * 1000 + batch)
    per = 1 if first else beam
    rows = per * batch
    logits = (4 * rs.randn(rows, V)).astype(np.float32)
    prev = (-5 * rs.rand(rows)).astype(np.float32)
    costs, keys = cuda.nth_element_logsoftmax(cuda.array(logits).t((per, 1, batch, V)), prev, batch, beam, n, first=first, suppress_word=suppress)
    ecosts, ekeys = expected_nbest(logits, prev, batch, per, n, suppress)
    costs, keys = costs.reshape(batch, n), keys.reshape(batch, n)
    # float rounding may swap neighbours whose totals agree to ~1e-6: compare the sets and the sorted costs, then the keys where costs are separated
    assert np.allclose(costs, ecosts, rtol=2e-6, atol=2e-5)
    for s in range(batch):
        gaps = np.abs(np.diff(ecosts[s])) > 1e-4
        if gaps.all():
            assert np.array_equal(keys[s], ekeys[s]), s
        else:
            assert set(keys[s]) <= set(np.lexsort((np.arange(per * V), -expected_total_row(logits, prev, batch, per, s, suppress)))[: n + 4] + s * per * V)


def expected_total_row(logits, prev, batch, per, s, suppress):
    V = logits.shape[-1]
    total = prev.astype(np.float64)[:, None] + logsoftmax(logits.astype(np.float64))
    if suppress >= 0:
        total[:, suppress] = np.finfo(np.float32).min
    return total.reshape(per, batch, V).transpose(1, 0, 2).reshape(batch, per * V)[s]


def test_fused_nbest_kernel_ties_go_to_the_lower_key(cuda):
    beam, batch, V, n = 3, 2, 9000, 4
    logits = np.zeros((beam * batch, V), dtype=np.float32)  # every word of every hypothesis ties
    prev = np.zeros(beam * batch, dtype=np.float32)
    costs, keys = cuda.nth_element_logsoftmax(cuda.array(logits).t((beam, 1, batch, V)), prev, batch, beam, n)
    assert np.array_equal(keys.reshape(batch, n), np.array([[0, 1, 2, 3], [beam * V, beam * V + 1, beam * V + 2, beam * V + 3]]))
    assert np.allclose(costs, -np.log(V), rtol=1e-6)


def test_range_nbest_kernel_matches_oracle(cuda, oracle):
    rs = np.random.RandomState(4)
    x = rs.randn(5 * 40000).astype(np.float32)
    x[7] = x[39999] = x[170000] = 11.0
    first = np.array([0, 40000, 40010, 100000, 200000])
    cum = np.array([0, 5, 8, 20, 21])
    c1, k1 = cuda.nth_element_ranges(cuda.array(x).t(), first, cum)
    c2, k2 = oracle.nth_element_ranges(oracle.array(x).t(), first, cum)
    assert np.array_equal(k1, k2) and np.array_equal(c1, c2)


def oracle_trained_pair(oracle, cuda, opts, mode, tmp_path):
    t = trained(oracle, opts)
    path = str(tmp_path / "model.npz")
    t.save(path)
    c = cuda.trainer(opts.replace("gemm-mode=0", "gemm-mode=%d" % mode))
    c.load(path)
    src, src_mask = t.get_batch(0)
    trg, trg_mask = t.get_batch(1)
    c.set_batch(src, src_mask, trg, trg_mask)
    return t, c


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("opts", [TRANSFORMER, S2S], ids=["transformer", "s2s"])
def test_search_on_gpu_returns_the_oracles_nbest_lists(cuda, oracle, opts, mode, tmp_path):
    t, c = oracle_trained_pair(oracle, cuda, opts, mode, tmp_path)
    for o in ("beam-size=4;normalize=0.6", "beam-size=1;allow-unk=true", "beam-size=3;beam-fused-nth=false", "beam-size=6;normalize=1;allow-unk=true"):
        exp, got = t.translate(o, n_best=3), c.translate(o, n_best=3)
        assert [[h[0] for h in s] for s in got] == [[h[0] for h in s] for s in exp], o
        assert np.allclose([h[1] for s in got for h in s], [h[1] for s in exp for h in s], rtol=1e-4), o
    t.close()
    c.close()


@pytest.mark.parametrize("opts", [TRANSFORMER, S2S], ids=["transformer", "s2s"])
def test_bf16_search_scores_equal_its_own_training_path(cuda, oracle, opts, tmp_path):
    t, c = oracle_trained_pair(oracle, cuda, opts, 4, tmp_path)
    src, src_mask = c.get_batch(0)
    fused = c.translate("beam-size=4;allow-unk=true;beam-fused-nth=true", n_best=2)
    plain = c.translate("beam-size=4;allow-unk=true;beam-fused-nth=false", n_best=2)
    assert [[h[0] for h in s] for s in fused] == [[h[0] for h in s] for s in plain]
    # against the oracle: bf16 products move the scores by ~1e-2; the best hypothesis rarely changes on these models
    exp = t.translate("beam-size=4;allow-unk=true", n_best=2)
    same = sum(f[0][0] == e[0][0] for f, e in zip(fused, exp))
    assert same >= len(exp) - 1
    for s, nbest in enumerate(fused):
        for words, score, raw in nbest:
            lp = sequence_logprobs(c, src[:, s:s + 1], src_mask[:, s:s + 1], words)
            assert abs(lp[np.arange(len(words)), words].sum() - raw) <= 3e-2 * max(1.0, abs(raw)), (s, words)
    t.close()
    c.close()


def test_copy_task_text_round_trip_on_gpu(cuda, tmp_path):
    from test_translator import write_copy_task

    lines = write_copy_task(tmp_path)
    # learn-rate 0.002 x 200 epochs: the copy model converges (CPU oracle: 17 of 17 lines and a cost of 0.003 for seeds
    # 3..6), so the count below does not move with the rounding-level differences between runs (atomics) and arithmetic
    # modes.  (60 epochs at 0.01, the CPU test's setting, leave a half-trained model: 12 - 16 of 17 over those seeds.)
    opts = ("type=transformer;dim-vocabs=16,16;dim-emb=32;transformer-heads=4;transformer-dim-ffn=64;enc-depth=1;dec-depth=1;tied-embeddings-all=true;"
            "workspace=128;graph-replay=true;gemm-mode=4;seed=3;learn-rate=0.002;clip-norm=1;label-smoothing=0")
    t = cuda.trainer(opts)
    vs = str(tmp_path / "vocab.src.yml")
    t.open_corpus(str(tmp_path / "train.src"), str(tmp_path / "train.trg"), vs, vs, "mini-batch=32;maxi-batch=4;seed=1")
    for _ in range(200):
        while t.next_corpus_batch():
            t.compute_gradients()
            t.update()
    (tmp_path / "test.src").write_text("\n".join(lines[:17]) + "\n")
    out = tmp_path / "test.out"
    assert t.translate_file(tmp_path / "test.src", vs, vs, out, "beam-size=3;normalize=0.6;mini-batch=5;maxi-batch=2") == 17
    got = out.read_text().splitlines()
    assert sum(g == e for g, e in zip(got, lines[:17])) >= 15, list(zip(got, lines[:17]))
    # training continues after decoding (graph plans, inference flag and staging are restored)
    while t.next_corpus_batch():
        t.compute_gradients()
        t.update()
    assert np.isfinite(t.cost())
    t.close()


@pytest.mark.parametrize("mode", [4, 3])
def test_transformer_base_dimensions_decode_and_timing(cuda, pkg, mode):
    opts = pkg.transformer_base_options(gemm_mode=mode)
    opts["graph-replay"] = "false"
    t = cuda.trainer(opts)
    t.next_synthetic_batch(64, 30, 30, padded=True)
    t.compute_gradients()
    t.update()  # parameters exist; random weights: hypotheses run to the length cap (3 x source length)
    cuda.synchronize()
    res = {}
    for o in ("beam-size=6;beam-fused-nth=true", "beam-size=6;beam-fused-nth=false"):
        t0 = time.time()
        out = t.translate(o, n_best=1)
        cuda.synchronize()
        res[o] = (time.time() - t0, out)
    (ta, a), (tb, b) = res["beam-size=6;beam-fused-nth=true"], res["beam-size=6;beam-fused-nth=false"]
    steps = max(len(h[0][0]) for h in a)
    words = sum(len(h[0][0]) for h in a)
    print("\n[decode transformer-base, 64 sentences x beam 6, mode %d] %d steps: fused selection %.2f s (%.1f ms/step, %.0f words/s), node sequence %.2f s"
          % (mode, steps, ta, 1e3 * ta / steps, words / ta, tb))
    assert len(a) == 64 and all(len(h) == 1 for h in a)
    agree = sum(x[0][0] == y[0][0] for x, y in zip(a, b))
    assert agree >= 60, agree  # random weights make near-ties common; the two selections normalise in different orders
    t.close()
