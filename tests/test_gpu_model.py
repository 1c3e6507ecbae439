"""Whole hot path on the GPU vs the CPU oracle: tape forward/backward over the
CUDA operators, clip + Adam, CUDA-graph replay.

The north-star tolerance: logits / loss within 1e-4 relative of the reference
arithmetic (here: the oracle) - checked in the fp32 SIMT mode and in the
bf16x3 tensor-core mode.  The throughput modes are bounded separately: tf32 (tcgen05
kind::tf32 on the fp32 tensors, 11-bit operands) at 4e-3, plain bf16 at 2e-2.
The CUDA Transformer uses the FUSED attention operator, the oracle builds the reference's
unfused node sequence - the comparison covers that fusion as well.
"""
import numpy as np
import pytest

from test_gpu_ops import close

pytestmark = pytest.mark.gpu

TRANSFORMER = ("type=transformer;dim-vocabs=200,220;dim-emb=64;transformer-heads=4;transformer-dim-ffn=128;"
               "enc-depth=2;dec-depth=2;workspace=256")
S2S_GRU = "type=s2s;dim-vocabs=200,220;dim-emb=32;dim-rnn=64;enc-depth=2;dec-depth=2;workspace=256"
N_PARAMS_BASE = 93_322_496  # Transformer-base, untied embeddings, V = 32000 per side (256 tensors)
S2S_LSTM = "type=s2s;dim-vocabs=200,220;dim-emb=32;dim-rnn=64;enc-cell=lstm;dec-cell=lstm;workspace=256"


def run_steps(lib, opts, mode, steps=3, batch=(8, 11, 13), padded=True, replay=False, keep=True):
    t = lib.trainer(opts + ";gemm-mode=%d;graph-replay=%s" % (mode, "true" if replay else "false"))
    out = {"costs": []}
    for s in range(steps):
        t.next_synthetic_batch(batch[0], batch[1], batch[2], padded=padded)
        t.compute_gradients(keep_logits=(keep and s == 0))
        if keep and s == 0:
            out["cost0"] = t.cost()
            out["logits"] = t.get_tensor("logits")
            out["grads"] = {n: t.get_tensor(n, grad=True) for n, _ in t.param_names()}
        t.update()
        out["costs"].append(t.cost())
    out["params"] = t.arena_numpy("params")
    out["stats"] = t.stats()
    t.close()
    return out


@pytest.mark.parametrize("opts", [TRANSFORMER, S2S_GRU, S2S_LSTM], ids=["transformer", "s2s-gru", "s2s-lstm"])
@pytest.mark.parametrize("mode,tol", [(0, 1e-4), (2, 1e-4), (1, 2e-2), (3, 4e-3), (4, 2e-2)], ids=["fp32", "bf16x3", "bf16", "tf32", "bf16-shadows"])
def test_step_matches_oracle(cuda, oracle, opts, mode, tol):
    exp = run_steps(oracle, opts, 0)
    got = run_steps(cuda, opts, mode)
    # loss and logits of the first step
    assert abs(got["cost0"] - exp["cost0"]) <= tol * abs(exp["cost0"]), (got["cost0"], exp["cost0"])
    close(got["logits"], exp["logits"], tol, "logits")
    # every parameter gradient (relative to that gradient's magnitude)
    # on the scale of the largest gradients (key-bias gradients are analytically zero)
    gtol = {0: 5e-4, 2: 5e-4, 1: 5e-2, 3: 1e-2, 4: 5e-2}[mode]
    gscale = max(float(np.abs(g).max()) for g in exp["grads"].values())
    for name, g in exp["grads"].items():
        scale = max(float(np.abs(g).max()), 1e-2 * gscale)
        err = float(np.abs(got["grads"][name].astype(np.float64) - g).max())
        assert err <= gtol * scale, "grad %s: %.3e (scale %.3e)" % (name, err, scale)
    # three clip+Adam updates: costs and the final flat parameter arena
    assert np.allclose(got["costs"], exp["costs"], rtol=tol * 3), (got["costs"], exp["costs"])
    if mode in (0, 2):
        # Adam moves every weight by ~lr per step whatever the gradient's size: weights with
        # analytically zero gradients (key biases) follow rounding noise, all others must agree
        diff = np.abs(got["params"] - exp["params"])
        assert diff.max() <= 2 * 3 * 1e-4 + 1e-5
        assert np.mean(diff > 2e-5) < 0.01 and np.median(diff) < 2e-6


@pytest.mark.parametrize("opts", [TRANSFORMER, S2S_GRU], ids=["transformer", "s2s-gru"])
@pytest.mark.parametrize("mode", [2, 4], ids=["bf16x3", "bf16-shadows"])
def test_graph_replay_equals_eager(cuda, opts, mode):
    """A captured+replayed step must produce what the eager tape produces (mode 4: the bf16 copy of the
    parameters is refreshed by Adam outside the captured graph)."""
    eager = run_steps(cuda, opts, mode, steps=6, padded=False, replay=False, keep=False)
    rep = run_steps(cuda, opts, mode, steps=6, padded=False, replay=True, keep=False)
    assert rep["stats"]["plans"] == 1 and rep["stats"]["replays"] >= 3, rep["stats"]
    assert np.allclose(rep["costs"], eager["costs"], rtol=2e-5), (rep["costs"], eager["costs"])
    diff = np.abs(rep["params"] - eager["params"])
    assert np.mean(diff > 2e-5) < 0.01 and np.median(diff) < 2e-6


@pytest.mark.parametrize("enc_type", ["bidirectional", "alternating", "bi-unidirectional"])
@pytest.mark.parametrize("mode", [2, 4], ids=["bf16x3", "bf16-shadows"])
def test_encoder_lanes_equal_single_stream(cuda, enc_type, mode):
    """The backward-direction stack of the s2s encoder runs as a lane of its own (own stream, ordered against the
    rest of the tape only where values / adjoints cross: ExpressionGraph::setLane).  Same results as the one-stream
    schedule (rnn-lanes=false), eagerly and through the captured graph, on padded batches (mask path)."""
    opts = "type=s2s;dim-vocabs=200,220;dim-emb=32;dim-rnn=64;enc-depth=3;dec-depth=2;enc-cell-depth=2;workspace=256;enc-type=" + enc_type
    for replay in (False, True):
        # (a plan is replayed only when the batch shape repeats: dense batches for the captured run, padded ones -
        # the mask path - for the eager run)
        one = run_steps(cuda, opts + ";rnn-lanes=false", mode, steps=6, replay=replay, keep=not replay, padded=not replay)
        two = run_steps(cuda, opts, mode, steps=6, replay=replay, keep=not replay, padded=not replay)
        assert np.allclose(two["costs"], one["costs"], rtol=2e-5), (replay, two["costs"], one["costs"])
        if not replay:
            close(two["logits"], one["logits"], 1e-6, "logits")
            gscale = max(float(np.abs(g).max()) for g in one["grads"].values())
            for name, g in one["grads"].items():  # bias gradients are sums of atomics: order noise only
                scale = max(float(np.abs(g).max()), 1e-2 * gscale)  # (analytically zero gradients, e.g. the attention bias, are rounding noise)
                assert float(np.abs(two["grads"][name] - g).max()) <= 2e-5 * scale, name
        diff = np.abs(two["params"] - one["params"])
        assert np.mean(diff > 2e-5) < 0.01 and np.median(diff) < 2e-6
        if replay:
            assert two["stats"]["plans"] == 1 and two["stats"]["replays"] >= 3, two["stats"]


@pytest.mark.parametrize("opts", [TRANSFORMER, S2S_GRU, S2S_LSTM], ids=["transformer", "s2s-gru", "s2s-lstm"])
@pytest.mark.parametrize("optimizer", ["adam", "sgd"])
def test_bf16_shadow_mode_equals_packed_bf16_model(cuda, opts, optimizer):
    """Mode 4 (bf16 shadows written by the producing kernels / Adam, TMA-direct) against mode 1 (every
    operand packed to bf16 by a separate pass): same arithmetic, so logits agree to accumulation noise and
    three updates stay together.  sgd: the optimizer does not write the bf16 parameter copy (refresh path)."""
    o = opts + ";optimizer=%s;learn-rate=%g" % (optimizer, 0.0001 if optimizer == "adam" else 0.05)
    a = run_steps(cuda, o, 4, steps=4)
    b = run_steps(cuda, o, 1, steps=4)
    close(a["logits"], b["logits"], 5e-5, "logits")
    assert abs(a["cost0"] - b["cost0"]) <= 1e-5 * abs(b["cost0"])
    assert np.allclose(a["costs"], b["costs"], rtol=5e-4), (a["costs"], b["costs"])


def test_fused_attention_equals_unfused_nodes(cuda):
    """Fused attention + fused residual/layer-norm on/off on the GPU, fp32 GEMM mode: same costs, logits, gradients."""
    outs = {}
    for fused in ("true", "false"):
        outs[fused] = run_steps(cuda, TRANSFORMER + ";transformer-fused-attention=%s;transformer-fused-residual-norm=%s" % (fused, fused), 0, steps=3, keep=True)
    a, b = outs["true"], outs["false"]
    assert np.allclose(a["costs"], b["costs"], rtol=2e-5), (a["costs"], b["costs"])
    close(a["logits"], b["logits"], 2e-5, "logits")
    gscale = max(float(np.abs(g).max()) for g in b["grads"].values())
    for name, g in b["grads"].items():
        err = float(np.abs(a["grads"][name].astype(np.float64) - g).max())
        assert err <= 1e-4 * max(float(np.abs(g).max()), 1e-2 * gscale), name


def test_replay_handles_changing_shapes(cuda):
    t = cuda.trainer(TRANSFORMER + ";gemm-mode=2;graph-replay=true")
    costs = []
    for s in range(9):
        shape = [(8, 11, 13), (8, 9, 7)][s % 2]   # two alternating shapes -> two plans
        t.next_synthetic_batch(shape[0], shape[1], shape[2], padded=False)
        t.compute_gradients()
        t.update()
        costs.append(t.cost())
    st = t.stats()
    assert st["plans"] == 2 and st["replays"] >= 4, st
    assert all(np.isfinite(costs))
    assert costs[-1] < costs[0]


def test_explicit_batch_equals_synthetic(cuda, oracle):
    """mrn_trainer_set_batch with host arrays in the CorpusBatch layout (time-major)."""
    rs = np.random.RandomState(0)
    B, Ts, Tt = 6, 9, 10
    src = rs.randint(2, 200, size=(Ts, B))
    trg = rs.randint(2, 220, size=(Tt, B))
    src[-1] = 0
    trg[-1] = 0
    sm, tm = np.ones((Ts, B), np.float32), np.ones((Tt, B), np.float32)
    sm[-2:, :2] = 0   # ragged: two sentences are two tokens shorter
    tm[-3:, 3:] = 0
    costs = []
    for lib in (cuda, oracle):
        t = lib.trainer(TRANSFORMER + ";gemm-mode=0;graph-replay=false")
        t.set_batch(src, sm, trg, tm)
        t.compute_gradients()
        costs.append(t.cost())
        assert t.batch_words() == (int(sm.sum()), int(sm.sum() + tm.sum()))
        t.close()
    assert abs(costs[0] - costs[1]) <= 1e-4 * abs(costs[1]), costs


def test_transformer_base_full_size_properties(cuda, pkg):
    """BASELINE.json config[1] at full size (64 x 50, V = 32000): properties that do
    not need the oracle at this size."""
    costs = {}
    for mode in (2, 1, 3, 4):
        t = cuda.trainer(pkg.transformer_base_options(gemm_mode=mode))
        cs = []
        for s in range(4):
            t.next_synthetic_batch(64, 50, 50, padded=False)
            t.compute_gradients()
            t.update()
            cs.append(t.cost())
        costs[mode] = cs
        names = t.param_names()
        assert sum(int(np.prod(s)) for _, s in names) == N_PARAMS_BASE, sum(int(np.prod(s)) for _, s in names)
        t.close()
    # untrained model on uniform random targets: cost per sentence ~ T * ln(V)
    assert abs(costs[2][0] - 50 * np.log(32000)) < 0.05 * 50 * np.log(32000), costs
    # bf16 vs bf16x3 on the same weights/batches
    assert np.allclose(costs[1], costs[2], rtol=2e-2), costs
    # tf32 vs bf16x3
    assert np.allclose(costs[3], costs[2], rtol=2e-3), costs
    # bf16 shadows (headline mode) vs packed bf16: same arithmetic
    assert np.allclose(costs[4], costs[1], rtol=1e-3), costs


def test_deep_gru_s2s_full_size_properties(cuda):
    """BASELINE.json config[2] at full size: deep GRU s2s (4+4, dim-emb 512, dim-rnn 1024, V = 32000),
    64 x 50 batches, tf32 GEMMs, graph replay.  ~10^4 kernels per step (GRU cells, Bahdanau attention,
    per-step GEMMs): checks the RNN rows of the scope table at the size the survey names."""
    opts = ("type=s2s;dim-vocabs=32000,32000;dim-emb=512;dim-rnn=1024;enc-depth=4;dec-depth=4;enc-cell=gru;dec-cell=gru;"
            "cost-type=ce-mean;learn-rate=0.0001;clip-norm=1;seed=1234;workspace=8192;gemm-mode=3;graph-replay=true")
    t = cuda.trainer(opts)
    costs = []
    for s in range(5):
        t.next_synthetic_batch(64, 50, 50, padded=False)
        t.compute_gradients()
        t.update()
        costs.append(t.cost())
    st = t.stats()
    t.close()
    assert all(np.isfinite(costs)), costs
    assert abs(costs[0] - 50 * np.log(32000)) < 0.05 * 50 * np.log(32000), costs
    assert costs[-1] < costs[0], costs
    assert st["plans"] == 1 and st["replays"] >= 2, st


def test_async_group_single_rank_equals_singleton(cuda, pkg):
    """AsyncGraphGroup with one rank (fetch from / push into its own master shard through the same
    lock + remote-Adam kernels the multi-GPU path uses) must reproduce SingletonGraph: same costs,
    same parameters (bias corrections computed with powf on the device: 1e-6 agreement)."""
    opts = TRANSFORMER + ";gemm-mode=0;graph-replay=true;learn-rate=0.001"
    costs = {"ref": [], "async": []}
    # one trainer after the other: parameter initialisation draws from a process-wide seed counter
    ref = cuda.trainer(opts)
    for s in range(5):
        ref.next_synthetic_batch(8, 11, 13, padded=True)
        ref.compute_gradients()
        ref.update()
        costs["ref"].append(ref.cost())
    pr = ref.arena_numpy("params")
    ref.close()
    a = pkg.AsyncTrainer(cuda, opts, 0)
    for s in range(5):
        a.trainer.next_synthetic_batch(8, 11, 13, padded=True)
        a.step()
        costs["async"].append(a.cost())
    a.fetch()  # replica <- master shards
    cuda.synchronize()
    pa = a.trainer.arena_numpy("params")
    a.trainer.close()
    assert np.allclose(costs["async"], costs["ref"], rtol=2e-5), costs
    diff = np.abs(pa - pr)
    assert np.mean(diff > 2e-5) < 0.01 and np.median(diff) < 2e-6, (diff.max(), np.mean(diff > 2e-5))


def test_async_group_optimizer_delay(cuda, pkg):
    """optimizer-delay tau = 2: gradients of two batches are accumulated and pushed once
    (reference graph_group_async.cu:186-215); costs stay finite and fall."""
    a = pkg.AsyncTrainer(cuda, TRANSFORMER + ";gemm-mode=2;graph-replay=true;learn-rate=0.002;optimizer-delay=2", 0)
    costs = []
    for s in range(12):
        a.trainer.next_synthetic_batch(8, 11, 13, padded=False)
        a.step()
        costs.append(a.cost())
    a.trainer.close()
    assert all(np.isfinite(costs)), costs
    assert np.mean(costs[-3:]) < np.mean(costs[:3]), costs


def test_dropout_draws_fresh_masks_on_every_replay(cuda):
    """Dropout inside a captured step: the seeds are baked into the CUDA graph, a device epoch
    counter bumped by the graph itself makes every replay draw new masks.  With frozen parameters
    (learn-rate 0) and the same batch, replayed costs must differ from step to step."""
    opts = TRANSFORMER + ";gemm-mode=0;graph-replay=true;transformer-dropout=0.3;learn-rate=0"
    t = cuda.trainer(opts)
    rs = np.random.RandomState(3)
    src, trg = rs.randint(2, 200, size=(9, 6)), rs.randint(2, 220, size=(10, 6))
    ones_s, ones_t = np.ones((9, 6), np.float32), np.ones((10, 6), np.float32)
    costs = []
    for s in range(6):
        t.set_batch(src, ones_s, trg, ones_t)
        t.compute_gradients()
        t.update()
        costs.append(t.cost())
    st = t.stats()
    t.close()
    assert st["replays"] >= 3, st
    assert len(set(np.round(costs[2:], 4))) == len(costs[2:]), costs   # replays: all different
    assert np.std(costs) < 0.05 * np.mean(costs), costs               # ... but the same model


def test_checkpoint_round_trip_between_gpu_and_oracle(cuda, oracle, tmp_path):
    """A model trained and saved on the GPU (reference .npz format, csrc/training/checkpoint.h) resumes on the GPU
    with its Adam state, and loads into the CPU oracle: same next-batch cost within the exact-mode tolerance."""
    opts = TRANSFORMER + ";gemm-mode=2;graph-replay=false;learn-rate=0.001"

    def steps(t, n, skip=0):
        for _ in range(skip):
            t.next_synthetic_batch(8, 11, 13, padded=True)
        out = []
        for _ in range(n):
            t.next_synthetic_batch(8, 11, 13, padded=True)
            t.compute_gradients()
            t.update()
            out.append(t.cost())
        return out

    ref = cuda.trainer(opts)
    want = steps(ref, 5)
    ref.close()
    a = cuda.trainer(opts)
    first = steps(a, 3)
    path = tmp_path / "gpu.npz"
    a.save(path, with_optimizer=True)
    a.close()
    b = cuda.trainer(opts)
    b.load(path, with_optimizer=True)
    rest = steps(b, 2, skip=3)
    b.close()
    assert np.allclose(first + rest, want, rtol=2e-5), (first + rest, want)
    o = oracle.trainer(opts.replace("gemm-mode=2", "gemm-mode=0"))
    o.load(path, with_optimizer=True)
    cpu = steps(o, 2, skip=3)
    o.close()
    assert np.allclose(cpu, want[3:], rtol=1e-4), (cpu, want[3:])


@pytest.mark.parametrize("padded", [False, True], ids=["dense", "padded"])
def test_host_running_ahead_does_not_mix_batches(cuda, padded):
    """bench.py never reads the cost inside its timed loop: the host enqueues steps far ahead of the device.  The
    pinned staging a replayed graph uploads from (and the eager tape's staging for new shapes) must not be refilled
    before the queued step has read it - otherwise steps train on the wrong batch or gather rows with torn indices
    (an illegal address in the padded bench of this round).  Same 40 updates with and without a host sync per step."""
    opts = TRANSFORMER + ";gemm-mode=0;graph-replay=true;learn-rate=0.002"
    out = []
    for sync in (True, False):
        t = cuda.trainer(opts)
        for s in range(40):
            shape = [(8, 11, 13), (8, 9, 7), (8, 12, 13)][s % 3] if padded else (8, 11, 13)
            t.next_synthetic_batch(shape[0], shape[1], shape[2], padded=padded)
            t.compute_gradients()
            t.update()
            if sync:
                t.cost()
        last = t.cost()
        out.append((last, t.arena_numpy("params")))
        t.close()
    assert abs(out[0][0] - out[1][0]) <= 1e-4 * abs(out[0][0]), (out[0][0], out[1][0])
    diff = np.abs(out[0][1] - out[1][1])
    assert np.mean(diff > 1e-4) < 0.01, (diff.max(), np.mean(diff > 1e-4))


def test_text_corpus_training_and_validation_on_gpu(cuda, oracle, tmp_path):
    """Text corpus -> prefetched mini-batches -> replayed training steps on the GPU, with cross-entropy validation in
    between (forward-only eager builds on the same graph): the validation result matches the oracle's on the same
    parameters (checkpoint round trip), and the replayed steps continue undisturbed."""
    import test_corpus as tc

    tc.write_corpus(tmp_path, n=96, max_len=8)
    vs, vt = str(tmp_path / "train.src") + ".yml", str(tmp_path / "train.trg") + ".yml"
    opts = tc.OPTS.replace("gemm-mode=0;graph-replay=false", "gemm-mode=2;graph-replay=true")
    t = cuda.trainer(opts)
    t.open_corpus(tmp_path / "train.src", tmp_path / "train.trg", options="mini-batch=16;maxi-batch=6;shuffle=false")
    costs = []
    for ep in range(4):
        while t.next_corpus_batch():
            t.compute_gradients()
            t.update()
            costs.append(t.cost())
        v = t.validate(tmp_path / "train.src", tmp_path / "train.trg", vs, vt, "valid-mini-batch=8")
        costs.append(v["metric"])
    assert t.stats()["replays"] >= 6, t.stats()
    path = tmp_path / "m.npz"
    t.save(path)
    t.close()
    o = oracle.trainer(tc.OPTS)
    o.load(path)
    ref = o.validate(tmp_path / "train.src", tmp_path / "train.trg", vs, vt, "valid-mini-batch=8")
    o.close()
    assert abs(v["metric"] - ref["metric"]) <= 1e-4 * abs(ref["metric"]), (v, ref)
    assert v["metric"] < costs[6], costs   # validation cost after four epochs below the first epoch's
