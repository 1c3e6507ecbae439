"""The persistent tile-loop GEMM (csrc/kernels/gemm.cu gGemmBf16Persistent: double-buffered TMEM accumulator,
column-sum warps) normally only takes products with >= 296 output tiles.  Here the bf16 product tests and the
model-level parity tests are re-run in a child process with MRN_GEMM_PERSIST_TILES=1, which sends EVERY eligible
product (N >= 128, one operand pair) through it - ragged tiles, single tiles, gated and column-sum variants."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("target", [
    "tests/test_gpu_gemm.py::test_bf16_shadow_mode_equals_packed_bf16",
    "tests/test_gpu_gemm.py::test_prod_swish_grad_nt",
    "tests/test_gpu_gemm.py::test_prod_affine",
    "tests/test_gpu_model.py::test_bf16_shadow_mode_equals_packed_bf16_model",
    "tests/test_gpu_model.py::test_graph_replay_equals_eager",
])
def test_persistent_kernel_everywhere(cuda, target):
    env = dict(os.environ, MRN_GEMM_PERSIST_TILES="1")
    r = subprocess.run([sys.executable, "-m", "pytest", target, "-m", "gpu", "-x", "-q"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


SNIPPET = r"""
import json, sys
sys.path.insert(0, %r)
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
out = {}
for name, opts in (("tiny", "type=transformer;dim-vocabs=200,224;dim-emb=64;transformer-heads=4;transformer-dim-ffn=128;enc-depth=2;dec-depth=2;workspace=256"),
                   ("base", None)):
    o = pkg.transformer_base_options(gemm_mode=4) if opts is None else opts + ";gemm-mode=4"
    for replay in ("false", "true"):
        oo = dict(o, **{"graph-replay": replay}) if isinstance(o, dict) else o + ";graph-replay=" + replay
        t = lib.trainer(oo)
        costs = []
        for s in range(4):
            t.next_synthetic_batch(*((64, 50, 50) if opts is None else (8, 11, 13)))
            t.compute_gradients(); t.update(); costs.append(t.cost())
        t.close()
        out[name + "-" + replay] = costs
print(json.dumps(out))
"""


def test_shadow_only_tensors_change_nothing(cuda):
    """Adjoints / activations whose only readers are products are written as bf16 copies only (kernels/shadow.h,
    shadowOnly).  With MRN_SHADOW_KEEP_FP32=1 every fp32 tensor is written as well: the arithmetic is the same, so
    the costs of four updates must agree to the run-to-run noise of the atomic reductions (split-K, column sums):
    exactly for the tiny model, 2e-5 for Transformer-base 64 x 50 - eager and replayed."""
    import json

    res = {}
    for keep in ("0", "1"):
        env = dict(os.environ)
        if keep == "1":
            env["MRN_SHADOW_KEEP_FP32"] = "1"
        r = subprocess.run([sys.executable, "-c", SNIPPET % ROOT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        res[keep] = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    import numpy as np

    for key in res["0"]:
        assert np.allclose(res["0"][key], res["1"][key], rtol=2e-5, atol=0), (key, res["0"][key], res["1"][key])
    assert res["0"]["tiny-false"] == res["1"]["tiny-false"], res
