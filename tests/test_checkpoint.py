"""Checkpoint / resume in the reference's .npz format (csrc/training/checkpoint.h, csrc/common/npz.h;
reference: src/graph/expression_graph.h:442-502, src/models/encdec.h:201-229), on the CPU oracle build
(the same host code as the product; tests/test_gpu_model.py repeats the round trip on the GPU)."""
import numpy as np
import pytest

OPTS = ("type=transformer;dim-vocabs=60,70;dim-emb=32;transformer-heads=4;transformer-dim-ffn=48;enc-depth=2;dec-depth=1;"
        "workspace=64;learn-rate=0.001;gemm-mode=0;graph-replay=false")
BATCH = (6, 7, 8)


def steps(t, n, skip=0):
    for _ in range(skip):
        t.next_synthetic_batch(*BATCH, padded=True)
    costs = []
    for _ in range(n):
        t.next_synthetic_batch(*BATCH, padded=True)
        t.compute_gradients()
        t.update()
        costs.append(t.cost())
    return costs


def test_saved_npz_is_readable_by_numpy_and_keeps_marian_names(oracle, tmp_path):
    t = oracle.trainer(OPTS)
    steps(t, 2)
    path = tmp_path / "model.npz"
    t.save(path)
    z = np.load(path)
    names = [n for n, _ in t.param_names()]
    assert sorted(z.files) == sorted(names + ["special:model.yml"])
    assert list(z.files)[:-1] == sorted(names)  # the reference writes in std::map (name) order, the yaml last
    for n, shape in t.param_names():
        assert z[n].dtype == np.float32 and z[n].shape == shape, n
        assert np.array_equal(z[n].ravel(), t.get_tensor(n)), n
    yml = bytes(z["special:model.yml"]).rstrip(b"\x00").decode()
    assert "type: transformer" in yml and "dim-emb: 32" in yml and "  - 60\n  - 70" in yml and "version:" in yml
    # names as the reference's Transformer creates them (src/models/transformer.h:194-261,318-350,384-449,495-662)
    for n in ("encoder_Wemb", "encoder_l1_self_Wq", "encoder_l1_self_bk", "encoder_l1_self_Wo_ln_scale", "encoder_l2_ffn_W1", "encoder_l2_ffn_ffn_ln_bias",
              "decoder_Wemb", "decoder_l1_context_Wv", "decoder_l1_context_Wo_ln_bias", "decoder_ff_logit_out_W", "decoder_ff_logit_out_b"):
        assert n in z.files, n
    t.close()


def test_resume_reproduces_the_uninterrupted_run(oracle, tmp_path):
    ref = oracle.trainer(OPTS)
    want = steps(ref, 5)
    ref.close()
    a = oracle.trainer(OPTS)
    first = steps(a, 3)
    path = tmp_path / "ckpt.npz"
    a.save(path, with_optimizer=True)
    a.close()
    b = oracle.trainer(OPTS)
    b.load(path, with_optimizer=True)
    rest = steps(b, 2, skip=3)  # the synthetic corpus restarts: skip the batches already seen
    b.close()
    assert np.allclose(first + rest, want, rtol=1e-6), (first + rest, want)
    # without the optimizer state the continuation differs (Adam restarts its moments)
    c = oracle.trainer(OPTS)
    c.load(path)
    other = steps(c, 2, skip=3)
    c.close()
    assert abs(other[0] - want[3]) < 1e-5 * abs(want[3])          # same parameters -> same cost of the next batch
    assert abs(other[1] - want[4]) > 1e-7 * abs(want[4])          # ... but a different update


def test_loads_a_file_written_by_numpy_with_reference_names(oracle, tmp_path):
    """A checkpoint as another Marian would write it: numpy.savez (ZIP64 members) of arrays under the reference's
    parameter names, vectors as 1-d arrays (the reference reshapes them to [1, n], expression_graph.h:458-462)."""
    t = oracle.trainer(OPTS)
    steps(t, 1)
    rs = np.random.RandomState(5)
    arrays = {}
    for n, shape in t.param_names():
        a = (0.05 * rs.standard_normal(shape)).astype(np.float32)
        arrays[n] = a.reshape(-1) if shape[0] == 1 and len(shape) == 2 and n.endswith(("_b", "_bq", "_bk", "_bv", "_bo", "_b1", "_b2")) else a
    t.close()
    path = tmp_path / "external.npz"
    np.savez(path, **arrays)
    u = oracle.trainer(OPTS)
    u.load(path)
    u.next_synthetic_batch(*BATCH, padded=True)
    u.compute_gradients()
    assert np.isfinite(u.cost())
    for n, _ in u.param_names():
        assert np.array_equal(u.get_tensor(n), arrays[n].ravel()), n
    u.close()


def test_load_rejects_a_different_model_and_compressed_archives(oracle, pkg, tmp_path):
    t = oracle.trainer(OPTS)
    steps(t, 1)
    path = tmp_path / "m.npz"
    t.save(path)
    arrays = {n: t.get_tensor(n).reshape(s) for n, s in t.param_names()}
    t.close()
    other = oracle.trainer(OPTS.replace("dim-emb=32", "dim-emb=64"))
    with pytest.raises(pkg.MarianError, match="dim-emb"):
        other.load(path)
    other.close()
    np.savez_compressed(tmp_path / "c.npz", **arrays)
    v = oracle.trainer(OPTS)
    with pytest.raises(pkg.MarianError, match="compressed"):
        v.load(tmp_path / "c.npz")
    v.close()
