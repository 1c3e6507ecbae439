"""Extracts the golden vectors of the reference's own unit tests into
tests/golden/reference_unit_tests.json.

Run in the build container (where /root/reference exists):
    python tests/golden/extract_reference_goldens.py
The GPU box has no /root/reference; it only sees the committed JSON.

Source files and the vectors taken from each (name in the C++ file -> case):
  src/tests/operator_tests.cpp:19-293   vC, vB2, smOut/lsmOut, vAdd/vMinus/vMult/vDiv,
                                        vT1/vA/vT3/vT4/vT5, vS1/vS2/9/776/vW, vO1..vO4, vLn
  src/tests/rnn_tests.cpp:32-250        vOutput, vContextSum1, vContextSum2
  src/tests/attention_tests.cpp:32-105  vAligned
  src/tests/graph_tests.cpp:16-55       zeros / ones / from_vector read-back
"""
import json
import os
import re

REF = "/root/reference/src/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_unit_tests.json")


def vectors(path):
    """All `std::vector<float> name({ ... })` literals of a file, in order."""
    text = open(path).read()
    text = re.sub(r"//[^\n]*", "", text)  # drop commented-out expectations
    out = {}
    for m in re.finditer(r"std::vector<float>\s+(\w+)\s*\(\s*\{([^}]*)\}\s*\)", text):
        vals = [float(x.strip().rstrip("f")) for x in m.group(2).replace("\n", " ").split(",") if x.strip()]
        out.setdefault(m.group(1), []).append(vals)
    return out


def main():
    op = vectors(os.path.join(REF, "operator_tests.cpp"))
    rnn = vectors(os.path.join(REF, "rnn_tests.cpp"))
    att = vectors(os.path.join(REF, "attention_tests.cpp"))

    # vA appears several times in operator_tests.cpp; the transpose section uses {1..8}
    vA8 = [v for v in op["vA"] if len(v) == 8][0]
    cases = {
        # exact comparisons in the reference (CHECK(values == v))
        "operator/dot": {"expected": op["vC"][0], "tol": "exact", "ref": "operator_tests.cpp:19-32"},
        "operator/scalar_mult": {"expected": op["vB2"][0], "tol": "exact", "ref": "operator_tests.cpp:34-46"},
        "operator/softmax": {"expected": op["smOut"][0] + op["lsmOut"][0], "tol": "approx", "ref": "operator_tests.cpp:48-78"},
        "operator/broadcast": {"expected": op["vAdd"][0] + op["vMinus"][0] + op["vMult"][0] + op["vDiv"][0], "tol": "approx", "ref": "operator_tests.cpp:80-119"},
        "operator/transpose": {"expected": op["vT1"][0] + vA8 + op["vT3"][0] + op["vT4"][0] + op["vT5"][0], "tol": "exact", "ref": "operator_tests.cpp:121-164"},
        "operator/reductions": {"expected": op["vS1"][0] + op["vS2"][0] + [9.0, 776.0] + op["vW"][0], "tol": "approx", "ref": "operator_tests.cpp:166-208"},
        "operator/concat": {"expected": op["vO1"][0] + op["vO2"][0] + op["vO3"][0] + op["vO4"][0], "tol": "exact", "ref": "operator_tests.cpp:210-262"},
        "operator/layer_norm": {"expected": op["vLn"][0], "tol": "approx", "ref": "operator_tests.cpp:264-293"},
        # 1 % tolerance in the reference (Approx(y).epsilon(0.01))
        "rnn/simple": {"expected": rnn["vOutput"][0], "tol": "1pct", "ref": "rnn_tests.cpp:32-68"},
        "rnn/s2s_encoder": {"expected": rnn["vContextSum1"][0] + rnn["vContextSum2"][0], "tol": "1pct", "ref": "rnn_tests.cpp:70-250"},
        "attention/context": {"expected": att["vAligned"][0], "tol": "1pct", "ref": "attention_tests.cpp:32-105"},
        "graph/param_init": {"expected": [0.0] * 6 + [1.0] * 6 + [1, 2, 3, 4, 5, 6], "tol": "exact", "ref": "graph_tests.cpp:16-55"},
    }
    with open(OUT, "w") as fh:
        json.dump(cases, fh, indent=1)
    print("wrote", OUT, {k: len(v["expected"]) for k, v in cases.items()})


if __name__ == "__main__":
    main()
