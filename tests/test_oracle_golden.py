"""Pins the CPU oracle to the reference: every golden vector of the reference's
own unit tests (src/tests/{operator,rnn,attention,graph}_tests.cpp) must be
reproduced by the oracle running the same graphs (tests/cpp/graph_golden.cpp)."""
import pytest

from conftest import check_golden

CASES = ["operator/dot", "operator/scalar_mult", "operator/softmax", "operator/broadcast", "operator/transpose",
         "operator/reductions", "operator/concat", "operator/layer_norm", "rnn/simple", "rnn/s2s_encoder",
         "attention/context", "graph/param_init"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_golden(oracle, goldens, case):
    check_golden(oracle.golden(case), goldens[case])


def test_all_golden_cases_covered(goldens):
    assert sorted(goldens) == sorted(CASES)
