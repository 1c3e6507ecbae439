"""Model code (csrc/models/transformer.h, s2s.h, layers/generic.cpp) is this repo's own writing of
the reference's models.  What must not drift: parameter names, shapes and CREATION ORDER (one seed
increment per tensor fixes the random-initialisation stream, and the names are the checkpoint keys
of the reference) and the training cost.  tests/golden/model_structure.json pins them for eight
configurations (pre/post-norm recipes, highway, tied embeddings, label smoothing, deep / alternating
/ bidirectional RNN stacks, LSTM)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOC = json.load(open(os.path.join(ROOT, "tests", "golden", "model_structure.json")))


@pytest.mark.parametrize("name", sorted(DOC["configs"]))
def test_parameters_and_cost_are_pinned(oracle, name):
    c = DOC["configs"][name]
    t = oracle.trainer(c["options"] + ";gemm-mode=0;graph-replay=false")
    t.next_synthetic_batch(5, 7, 8, padded=True)
    t.compute_gradients()
    t.update()
    params = [[n, list(s)] for n, s in t.param_names()]
    cost = t.cost()
    t.close()
    assert params == [[n, list(s)] for n, s in c["params"]]
    assert abs(cost - c["cost_after_one_update"]) <= 1e-5 * abs(c["cost_after_one_update"])
