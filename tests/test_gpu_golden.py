"""The CUDA path against the reference's golden vectors (same driver source as
the oracle test, linked against libmarian_b200.so)."""
import pytest

from conftest import check_golden
from test_oracle_golden import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES)
def test_cuda_reproduces_reference_golden(cuda, goldens, case):
    check_golden(cuda.golden(case), goldens[case])
