"""Text corpus -> vocabularies -> length-sorted mini-batches (csrc/data/corpus.h) against a Python restatement of
the reference's rules (src/data/vocab.cpp, src/data/corpus.cpp:146-187, src/data/batch_generator.h:39-160), on the
CPU oracle build (same host code as the product)."""
import collections

import numpy as np
import pytest

OPTS = ("type=transformer;dim-vocabs=64,64;dim-emb=32;transformer-heads=4;transformer-dim-ffn=48;enc-depth=1;dec-depth=1;"
        "workspace=64;learn-rate=0.003;gemm-mode=0;graph-replay=false")


def write_corpus(tmp_path, n=137, seed=3, vocab=40, max_len=14):
    rs = np.random.RandomState(seed)
    words = ["w%d" % i for i in range(vocab)]
    # Zipf-ish frequencies so that the frequency order is unambiguous for the head of the vocabulary
    p = 1.0 / np.arange(1, vocab + 1) ** 1.3
    p /= p.sum()
    src, trg = [], []
    for _ in range(n):
        ls, lt = rs.randint(1, max_len), rs.randint(1, max_len)
        s = list(rs.choice(words, size=ls, p=p))
        src.append(" ".join(s))
        trg.append(" ".join(reversed(s[:lt])) if lt <= ls else " ".join(s + list(rs.choice(words, size=lt - ls, p=p))))
    (tmp_path / "train.src").write_text("\n".join(src) + "\n")
    (tmp_path / "train.trg").write_text("\n".join(trg) + "\n")
    return src, trg


def read_vocab(path):
    v = {}
    for line in open(path):
        k, i = line.rstrip("\n").rsplit(":", 1)
        v[k.strip().strip('"')] = int(i)
    return v


def epoch(t):
    out = []
    while t.next_corpus_batch():
        out.append((t.get_batch(0), t.get_batch(1)))
    return out


def test_vocabulary_is_created_by_falling_frequency(oracle, tmp_path):
    src, trg = write_corpus(tmp_path)
    t = oracle.trainer(OPTS)
    t.open_corpus(tmp_path / "train.src", tmp_path / "train.trg", options="mini-batch=8;maxi-batch=4")
    v = read_vocab(str(tmp_path / "train.src") + ".yml")
    assert v["</s>"] == 0 and v["<unk>"] == 1
    counts = collections.Counter(w for line in src for w in line.split())
    by_id = sorted((i, w) for w, i in v.items() if i >= 2)
    freqs = [counts[w] for _, w in by_id]
    assert freqs == sorted(freqs, reverse=True) and [i for i, _ in by_id] == list(range(2, 2 + len(by_id)))
    t.close()


@pytest.mark.parametrize("sort", ["trg", "src"])
def test_epoch_covers_every_sentence_once_in_length_sorted_batches(oracle, tmp_path, sort):
    src, trg = write_corpus(tmp_path)
    t = oracle.trainer(OPTS)
    t.open_corpus(tmp_path / "train.src", tmp_path / "train.trg", options="mini-batch=8;maxi-batch=4;maxi-batch-sort=%s;shuffle=false;max-length=12" % sort)
    vs, vt = read_vocab(str(tmp_path / "train.src") + ".yml"), read_vocab(str(tmp_path / "train.trg") + ".yml")
    batches = epoch(t)
    # expected sentence set: tuples with both sides <= max-length INCLUDING the appended </s> (corpus.cpp:180-184)
    want = collections.Counter()
    for s, g in zip(src, trg):
        si = [vs[w] for w in s.split()] + [0]
        ti = [vt[w] for w in g.split()] + [0]
        if len(si) <= 12 and len(ti) <= 12:
            want[(tuple(si), tuple(ti))] += 1
    got = collections.Counter()
    for (si, sm), (ti, tm) in batches:
        B = si.shape[1]
        assert B <= 8 and ti.shape[1] == B
        for b in range(B):
            ls, lt = int(sm[:, b].sum()), int(tm[:, b].sum())
            # mask = prefix of ones, padding index 0, last real token = </s>
            assert np.all(sm[:ls, b] == 1) and np.all(sm[ls:, b] == 0) and si[ls - 1, b] == 0 and np.all(si[ls:, b] == 0)
            got[(tuple(si[:ls, b]), tuple(ti[:lt, b]))] += 1
        assert si.shape[0] == int(sm.sum(0).max()) and ti.shape[0] == int(tm.sum(0).max())   # padded to the batch maximum
    assert got == want
    # maxi-batches of 32 sentences, popped longest first: within a maxi-batch the batch widths do not increase
    key = 1 if sort == "trg" else 0
    widths = [b[key][0].shape[0] for b in batches]
    for m in range(0, len(widths), 4):
        chunk = widths[m:m + 4]
        assert chunk == sorted(chunk, reverse=True), (m, widths)
    # a second epoch (after the end-of-epoch marker) yields the same batches when not shuffling
    again = epoch(t)
    assert len(again) == len(batches) and all(np.array_equal(a[0][0], b[0][0]) for a, b in zip(again, batches))
    t.close()


def test_mini_batch_words_and_shuffling(oracle, tmp_path):
    write_corpus(tmp_path)
    t = oracle.trainer(OPTS)
    t.open_corpus(tmp_path / "train.src", tmp_path / "train.trg", options="mini-batch=64;maxi-batch=2;mini-batch-words=40;shuffle=false")
    small = 0
    for (si, sm), _ in epoch(t):
        words = int(sm.sum())
        longest = int(sm.sum(0).max())
        assert words <= 40 + longest   # a batch is closed by the sentence that pushes it past the limit
        small += words <= 40
    assert small <= 2   # only the remainder of each of the two maxi-batches (128 sentences each) may stay below it
    t.close()
    a = oracle.trainer(OPTS)
    a.open_corpus(tmp_path / "train.src", tmp_path / "train.trg", options="mini-batch=8;maxi-batch=4;shuffle=true;seed=7")
    e1, e2 = epoch(a), epoch(a)
    assert sum(b[0][0].shape[1] for b in e1) == sum(b[0][0].shape[1] for b in e2)
    assert any(x[0][0].shape != y[0][0].shape or not np.array_equal(x[0][0], y[0][0]) for x, y in zip(e1, e2))  # reshuffled every epoch
    a.close()


def test_training_on_text_batches_learns(oracle, tmp_path):
    write_corpus(tmp_path, n=96, max_len=8)
    t = oracle.trainer(OPTS)
    t.open_corpus(tmp_path / "train.src", tmp_path / "train.trg", options="mini-batch=16;maxi-batch=6")
    costs = []
    for ep in range(6):
        while t.next_corpus_batch():
            t.compute_gradients()
            t.update()
            costs.append(t.cost() / max(1, t.batch_words()[1]))
    assert np.mean(costs[-6:]) < 0.8 * np.mean(costs[:6]), (costs[:6], costs[-6:])
    t.close()


def test_cross_entropy_validation(oracle, tmp_path):
    """CrossEntropyValidator (csrc/training/validator.h; reference src/training/validator.h:108-176): forward-only
    passes over a held-out corpus on the trainer's parameters, reported per cost-type; training goes on unchanged."""
    write_corpus(tmp_path, n=96, max_len=8)
    dev = tmp_path / "dev"
    dev.mkdir()
    write_corpus(dev, n=40, seed=11, max_len=8)
    vs, vt = str(tmp_path / "train.src") + ".yml", str(tmp_path / "train.trg") + ".yml"

    def train(t, epochs):
        for _ in range(epochs):
            while t.next_corpus_batch():
                t.compute_gradients()
                t.update()
                t.cost()

    t = oracle.trainer(OPTS)
    t.open_corpus(tmp_path / "train.src", tmp_path / "train.trg", options="mini-batch=16;maxi-batch=6;shuffle=false")
    train(t, 1)
    before = t.validate(dev / "train.src", dev / "train.trg", vs, vt, "valid-mini-batch=8")
    p0 = t.arena_numpy("params")
    again = t.validate(dev / "train.src", dev / "train.trg", vs, vt, "valid-mini-batch=8")
    assert np.array_equal(p0, t.arena_numpy("params"))                      # validation does not touch the parameters
    assert again["metric"] == before["metric"]
    assert before["sentences"] == 40 and before["target_words"] > 40
    assert abs(before["metric"] - before["cost_sum"] / 40) < 1e-4 * abs(before["metric"])   # ce-mean: cost per sentence
    words = t.validate(dev / "train.src", dev / "train.trg", vs, vt, "cost-type=ce-mean-words")
    ppl = t.validate(dev / "train.src", dev / "train.trg", vs, vt, "cost-type=perplexity")
    assert abs(words["metric"] - words["cost_sum"] / words["target_words"]) < 1e-5 * words["metric"]
    assert abs(ppl["metric"] - np.exp(words["metric"])) < 1e-3 * ppl["metric"]
    # batching does not change the summed cost (up to fp32 summation order)
    other = t.validate(dev / "train.src", dev / "train.trg", vs, vt, "valid-mini-batch=3;cost-type=ce-sum")
    assert abs(other["metric"] - before["cost_sum"]) < 2e-5 * before["cost_sum"]
    train(t, 5)
    after = t.validate(dev / "train.src", dev / "train.trg", vs, vt, "valid-mini-batch=8")
    assert after["metric"] < 0.9 * before["metric"], (before, after)            # the copy-like task generalises to the dev set
    t.close()
