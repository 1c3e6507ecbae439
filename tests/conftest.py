import json
import os
import sys

import numpy as np
import pytest

# The oracle parallelises with OpenMP; on a 128-core GPU host the fork/join cost of tiny
# loops dominates the small-model parity tests, so the test suite caps the team size.
os.environ.setdefault("OMP_NUM_THREADS", "16")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (TEST INFRASTRUCTURE): oracle/_build/libmarian_oracle.so."""
    if not os.path.exists(graft.ORACLE_LIB):
        graft._load_build().build_oracle()
    return graft.load_oracle()


@pytest.fixture(scope="session")
def cuda(pkg):
    """The product library on a real GPU; fails loudly (no fallback) if it cannot run."""
    lib = pkg.load()
    assert lib.backend == "cuda"
    a = lib.array(np.arange(4, dtype=np.float32))
    assert np.array_equal(a.numpy(), np.arange(4, dtype=np.float32))
    return lib


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_unit_tests.json")) as fh:
        return json.load(fh)


def check_golden(values, case):
    """Tolerance classes of the reference's Catch tests."""
    exp = np.array(case["expected"], dtype=np.float64)
    got = np.array(values, dtype=np.float64)
    assert got.shape == exp.shape
    if case["tol"] == "exact":
        assert np.array_equal(got, exp), (got, exp)
    elif case["tol"] == "approx":
        # Catch Approx default: |x - y| < eps * (1 + |y|), eps = 100 * FLT_EPSILON; the golden
        # literals themselves carry ~6 significant digits
        assert np.all(np.abs(got - exp) <= 1.2e-5 * (1 + np.abs(exp)) + 5e-6 * np.abs(exp) + 1e-6), np.abs(got - exp).max()
    else:  # "1pct": Approx(y).epsilon(0.01)
        assert np.all(np.abs(got - exp) <= 0.01 * (1 + np.abs(exp))), np.abs(got - exp).max()
