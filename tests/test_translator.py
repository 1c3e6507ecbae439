"""Beam-search decoding (csrc/translator/beam_search.h, csrc/translator/translator.h) on the CPU oracle build - the same
host code as the product, over the CPU operator layer.

What pins it (the reference has no decoding tests):
  * the n-best selection against numpy (ranges, fused log-softmax variant, ties, suppressed word, first step);
  * the incremental decoder path (state selection per hypothesis, layer-input caches, positions, single-step RNN) against the
    FULL-SEQUENCE training path of the same model: the score of every returned hypothesis equals minus the cross-entropy the
    training graph assigns to it, and with beam 1 every word is the arg-max continuation of its prefix;
  * the search conventions of src/translator/beam_search.h:91-225 against a Python restatement that drives the model
    through the training path only (exhaustive re-scoring of all prefixes, tiny vocabulary);
  * the reference's node sequence (logsoftmax + add + transpose + n-best over ranges) and the fused operator give identical results;
  * a copy task learnt from a text corpus is translated back correctly, in corpus order (translator.h).
"""
import numpy as np
import pytest

TRANSFORMER = ("type=transformer;dim-vocabs=24,24;dim-emb=32;transformer-heads=4;transformer-dim-ffn=64;enc-depth=2;dec-depth=2;"
               "workspace=128;graph-replay=false;gemm-mode=0;seed=7;learn-rate=0.003;label-smoothing=0")
S2S = ("type=s2s;dim-vocabs=24,24;dim-emb=24;dim-rnn=32;enc-depth=1;dec-depth=1;enc-cell=gru;dec-cell=gru;enc-cell-depth=1;dec-cell-base-depth=2;"
       "workspace=128;graph-replay=false;gemm-mode=0;seed=9;learn-rate=0.003;label-smoothing=0")


def logsoftmax(x):
    m = x.max(axis=-1, keepdims=True)
    return (x - m) - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def trained(lib, opts, steps=150, batch=6, ls=5, lt=6):
    t = lib.trainer(opts)
    for _ in range(steps):
        t.next_synthetic_batch(batch, ls, lt, padded=True)
        t.compute_gradients()
        t.update()
    t.next_synthetic_batch(batch, ls, lt, padded=True)
    return t


def sequence_logprobs(t, src, src_mask, words):
    """log p(word_k | prefix) for one target sequence through the training graph (full-sequence path)."""
    trg = np.array(words, dtype=np.int64).reshape(-1, 1)
    t.set_batch(src, src_mask, trg, np.ones(trg.shape, dtype=np.float32))
    t.compute_gradients(keep_logits=True)
    logits = t.get_tensor("logits").reshape(len(words), -1)
    return logsoftmax(logits.astype(np.float64))


# ------------------------------------------------------------------------------------------------------------------
# n-best selection operators
# ------------------------------------------------------------------------------------------------------------------
def test_nth_element_ranges_matches_numpy(oracle):
    rs = np.random.RandomState(0)
    x = rs.randn(7 * 300).astype(np.float32)
    x[5] = x[17] = x[211] = 9.0  # ties: lower index first
    first = np.array([0, 300, 900, 1000, 2100])
    cum = np.array([0, 4, 8, 9, 15])
    costs, keys = oracle.nth_element_ranges(oracle.array(x).t(), first, cum)
    for r in range(4):
        seg = x[first[r]:first[r + 1]]
        order = np.lexsort((np.arange(len(seg)), -seg))[: cum[r + 1] - cum[r]] + first[r]
        assert np.array_equal(keys[cum[r]:cum[r + 1]], order)
        assert np.array_equal(costs[cum[r]:cum[r + 1]], x[order])


@pytest.mark.parametrize("first", [False, True])
@pytest.mark.parametrize("suppress", [-1, 1])
def test_nth_element_logsoftmax_matches_numpy(oracle, first, suppress):
    rs = np.random.RandomState(1)
    beam, batch, V, n = 3, 4, 37, 3
    rows = (1 if first else beam) * batch
    logits = (3 * rs.randn(rows, V)).astype(np.float32)
    prev = rs.randn(rows).astype(np.float32)
    costs, keys = oracle.nth_element_logsoftmax(oracle.array(logits).t((1 if first else beam, 1, batch, V)), prev, batch, beam, n, first=first, suppress_word=suppress)
    total = prev[:, None] + logsoftmax(logits.astype(np.float64))
    if suppress >= 0:
        total[:, suppress] = np.finfo(np.float32).min
    per = 1 if first else beam
    total = total.reshape(per, batch, V).transpose(1, 0, 2).reshape(batch, per * V)
    for s in range(batch):
        order = np.lexsort((np.arange(per * V), -total[s]))[:n]
        assert np.array_equal(keys[s * n:(s + 1) * n], order + s * per * V)
        assert np.allclose(costs[s * n:(s + 1) * n], total[s][order], rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------------------------
# decoder path against the training path
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("opts", [TRANSFORMER, S2S], ids=["transformer", "s2s"])
def test_hypothesis_scores_equal_training_path_cross_entropy(oracle, opts):
    t = trained(oracle, opts)
    src, src_mask = t.get_batch(0)
    results = t.translate("beam-size=3;normalize=0;allow-unk=true", n_best=3)
    assert len(results) == src.shape[1]
    checked = 0
    for s, nbest in enumerate(results):
        assert 1 <= len(nbest) <= 3
        scores = [h[1] for h in nbest]
        assert scores == sorted(scores, reverse=True)
        assert len({tuple(h[0]) for h in nbest}) == len(nbest)
        for words, score, raw in nbest:
            assert score == raw  # normalize=0
            assert len(words) <= 3 * src.shape[0]
            lp = sequence_logprobs(t, src[:, s:s + 1], src_mask[:, s:s + 1], words)
            assert abs(lp[np.arange(len(words)), words].sum() - raw) <= 2e-4 * max(1.0, abs(raw)), (s, words)
            checked += 1
    assert checked >= len(results)
    t.close()


@pytest.mark.parametrize("opts", [TRANSFORMER, S2S], ids=["transformer", "s2s"])
def test_beam_one_is_greedy_decoding(oracle, opts):
    t = trained(oracle, opts)
    src, src_mask = t.get_batch(0)
    results = t.translate("beam-size=1;allow-unk=true", n_best=1)
    for s, nbest in enumerate(results):
        words = nbest[0][0]
        lp = sequence_logprobs(t, src[:, s:s + 1], src_mask[:, s:s + 1], words)
        assert np.array_equal(lp.argmax(axis=1), np.array(words)), (s, words)
        # a finished hypothesis ends with </s> and has none before
        assert 0 not in words[:-1]
    t.close()


def test_unknown_word_is_suppressed_by_default(oracle):
    t = trained(oracle, TRANSFORMER)
    for nbest in t.translate("beam-size=4", n_best=4):
        for words, _, _ in nbest:
            assert 1 not in words
    t.close()


def test_length_normalisation_ranks_by_cost_over_length(oracle):
    t = trained(oracle, TRANSFORMER)
    for nbest in t.translate("beam-size=4;normalize=1;allow-unk=true", n_best=4):
        for words, score, raw in nbest:
            # history.h: cost / (steps so far)^alpha; a hypothesis added at step k has k words
            assert abs(score - raw / len(words)) <= 1e-5 * max(1.0, abs(score))
        assert [h[1] for h in nbest] == sorted((h[1] for h in nbest), reverse=True)
    t.close()


@pytest.mark.parametrize("opts", [TRANSFORMER, S2S], ids=["transformer", "s2s"])
def test_fused_selection_equals_reference_node_sequence(oracle, opts):
    t = trained(oracle, opts)
    for o in ("beam-size=5;normalize=0.6", "beam-size=2;allow-unk=true", "beam-size=1"):
        a = t.translate(o + ";beam-fused-nth=true", n_best=2)
        b = t.translate(o + ";beam-fused-nth=false", n_best=2)
        assert [[h[0] for h in s] for s in a] == [[h[0] for h in s] for s in b]
        assert np.allclose([h[1] for s in a for h in s], [h[1] for s in b for h in s], rtol=1e-5)
    t.close()


# ------------------------------------------------------------------------------------------------------------------
# the search loop against a Python restatement that only uses the training path
# ------------------------------------------------------------------------------------------------------------------
def python_beam_search(t, src, src_mask, beam, max_steps, allow_unk=True):
    """src/translator/beam_search.h:91-225 for ONE sentence, re-scoring every prefix through the full-sequence training
    path: start with `beam` empty hypotheses of which only the first competes; every step keep the `width` best
    continuations over all live hypotheses; hypotheses ending in </s> (word 0) are recorded and leave the beam; the width of
    the next step is the number of live hypotheses; stop when none is left or after max_steps (then everything is recorded)."""
    live, finished, steps = [([], 0.0)], [], 1
    width = beam
    first = True
    while True:
        cands = []
        for hi, (words, cost) in enumerate(live):
            lp = sequence_logprobs(t, src, src_mask, words + [0])[len(words)].astype(np.float32)
            if not allow_unk:
                lp[1] = np.finfo(np.float32).min
            for w in range(len(lp)):
                cands.append((np.float32(cost) + lp[w], hi, w))
        cands.sort(key=lambda c: (-c[0], c[1], c[2]))
        keep = len(live) if not first else beam
        chosen = cands[:width][:keep]
        cut = steps >= max_steps
        nxt = [(live[hi][0] + [w], float(c)) for c, hi, w in chosen]
        alive = [h for h in nxt if h[0][-1] != 0]
        last = not alive or cut
        for h in nxt:
            if h[0][-1] == 0 or last:
                finished.append((h[1], steps, h[0]))
        steps += 1
        live = alive
        width = beam if first else len(live)  # the reference narrows the beam only from the second step on
        first = False
        if not live or cut:
            break
    finished.sort(key=lambda f: -f[0])
    return finished


def test_search_loop_matches_python_restatement(oracle):
    t = trained(oracle, TRANSFORMER)
    src, src_mask = t.get_batch(0)
    got = t.translate("beam-size=3;normalize=0;allow-unk=true", n_best=3)
    for s in range(src.shape[1]):
        exp = python_beam_search(t, src[:, s:s + 1], src_mask[:, s:s + 1], 3, 3 * src.shape[0])
        # restore the batch the trainer translates (sequence_logprobs replaced it)
        for (words, score, _), (escore, _, ewords) in zip(got[s], exp):
            assert words == ewords, (s, got[s], exp[:3])
            assert abs(score - escore) <= 2e-4 * max(1.0, abs(escore))
    t.close()


# ------------------------------------------------------------------------------------------------------------------
# text in, text out
# ------------------------------------------------------------------------------------------------------------------
def write_copy_task(tmp_path, n=96, vocab=12, seed=5):
    rs = np.random.RandomState(seed)
    words = ["w%d" % i for i in range(vocab)]
    lines = [" ".join(rs.choice(words, size=rs.randint(2, 6))) for _ in range(n)]
    (tmp_path / "train.src").write_text("\n".join(lines) + "\n")
    (tmp_path / "train.trg").write_text("\n".join(lines) + "\n")
    return lines


def test_copy_task_round_trip_through_text_files(oracle, tmp_path):
    lines = write_copy_task(tmp_path)
    opts = ("type=transformer;dim-vocabs=16,16;dim-emb=32;transformer-heads=4;transformer-dim-ffn=64;enc-depth=1;dec-depth=1;tied-embeddings-all=true;"
            "workspace=128;graph-replay=false;gemm-mode=0;seed=3;learn-rate=0.01;clip-norm=1;label-smoothing=0")
    t = oracle.trainer(opts)
    src, trg = str(tmp_path / "train.src"), str(tmp_path / "train.trg")
    vs, vt = str(tmp_path / "vocab.src.yml"), str(tmp_path / "vocab.trg.yml")
    t.open_corpus(src, trg, vs, vs, "mini-batch=32;maxi-batch=4;seed=1")
    for _ in range(60):
        while t.next_corpus_batch():
            t.compute_gradients()
            t.update()
    test_lines = lines[:17]
    (tmp_path / "test.src").write_text("\n".join(test_lines) + "\n")
    out = tmp_path / "test.out"
    n = t.translate_file(tmp_path / "test.src", vs, vs, out, "beam-size=3;normalize=0.6;mini-batch=5;maxi-batch=2")
    assert n == len(test_lines)
    got = out.read_text().splitlines()
    assert len(got) == len(test_lines)
    same = sum(g == e for g, e in zip(got, test_lines))
    assert same >= len(test_lines) - 2, list(zip(got, test_lines))
    # n-best output format of the reference's output collector
    nb = tmp_path / "test.nbest"
    t.translate_file(tmp_path / "test.src", vs, vs, nb, "beam-size=2;n-best=true")
    rows = nb.read_text().splitlines()
    assert len(rows) >= len(test_lines)
    ident, text, feat, total = rows[0].split(" ||| ")
    assert ident == "0" and feat.startswith("F0= ") and float(total) <= 0
    t.close()
