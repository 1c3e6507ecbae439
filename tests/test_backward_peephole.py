"""Host logic of the backward-sweep peepholes (graph/node.h fuseBackward, AffineNodeOp):
projections of the same input that follow each other in the sweep hand their input gradients to ONE
grouped product.  On the CPU oracle the grouped product is the chain of accumulating products, issued
in the same order as the ungrouped sweep - so costs, gradients and updated parameters must be
bit-identical with the peephole on and off.  This pins the selection logic (which nodes are grouped,
that none is counted twice or dropped) without a GPU; the GPU tiers compare the grouped tensor-core
launch itself with the oracle.  The toggle is read once per process, hence the subprocesses."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import hashlib, json, sys
import numpy as np
sys.path.insert(0, %(root)r)
import __graft_entry__ as graft
lib = graft.load_oracle()
t = lib.trainer(%(opts)r + ";gemm-mode=0;graph-replay=false")
out = {"costs": []}
for s in range(2):
    t.next_synthetic_batch(4, 7, 9, padded=True)
    t.compute_gradients(keep_logits=False)
    if s == 0:
        out["grads"] = {n: hashlib.sha256(np.ascontiguousarray(t.get_tensor(n, grad=True)).tobytes()).hexdigest() for n, _ in t.param_names()}
    t.update()
    out["costs"].append(float(t.cost()))
out["params"] = hashlib.sha256(np.ascontiguousarray(t.arena_numpy("params")).tobytes()).hexdigest()
t.close()
print("RESULT " + json.dumps(out))
"""

MODELS = {
    "transformer": "type=transformer;dim-vocabs=60,70;dim-emb=32;transformer-heads=4;transformer-dim-ffn=64;enc-depth=2;dec-depth=2;workspace=128",
    "s2s-gru": "type=s2s;dim-vocabs=60,70;dim-emb=16;dim-rnn=32;enc-depth=1;dec-depth=1;workspace=128",
}


def run(opts, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    env.setdefault("OMP_NUM_THREADS", "4")
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "opts": opts}], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    out["trace"] = [l for l in p.stderr.splitlines() if l.startswith("[peephole]")]
    return out


@pytest.mark.parametrize("model", sorted(MODELS))
def test_grouped_input_gradients_equal_ungrouped(model):
    on = run(MODELS[model], {"MRN_PEEPHOLE_TRACE": "1"})
    off = run(MODELS[model], {"MRN_NO_GROUPED_DX": "1", "MRN_PEEPHOLE_TRACE": "1"})
    assert not off["trace"]
    if model == "transformer":
        # per step: q/k/v of the 2 + 2 self-attention blocks, key/value of the 2 cross-attention blocks
        per_step = len(on["trace"]) // 2
        assert sum("of 3 projections" in l for l in on["trace"]) == 2 * 4 and sum("of 2 projections" in l for l in on["trace"]) == 2 * 2, on["trace"][:per_step]
    assert on["costs"] == off["costs"], (on["costs"], off["costs"])
    assert np.all(np.isfinite(on["costs"]))
    diff = [n for n in on["grads"] if on["grads"][n] != off["grads"][n]]
    assert not diff, "gradients differ with the peephole on/off: %s" % diff[:5]
    assert on["params"] == off["params"]
