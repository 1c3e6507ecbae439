"""Regenerates tests/golden/model_structure.json from the CPU oracle (the option strings are read
back from the existing file).  Only run this deliberately: the file pins parameter names, shapes and
creation order (= the initialisation stream) of the model code across refactors."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

path = os.path.join(ROOT, "tests", "golden", "model_structure.json")
doc = json.load(open(path))
orc = graft.load_oracle()
for name, c in doc["configs"].items():
    t = orc.trainer(c["options"] + ";gemm-mode=0;graph-replay=false")
    t.next_synthetic_batch(5, 7, 8, padded=True)
    t.compute_gradients()
    t.update()
    c["params"] = [(n, list(s)) for n, s in t.param_names()]
    c["cost_after_one_update"] = t.cost()
    t.close()
json.dump(doc, open(path, "w"), indent=0)
