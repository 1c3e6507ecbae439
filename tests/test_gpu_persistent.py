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
