"""Runs a few EAGER (no graph replay) Transformer-base training steps so that ncu can list
every kernel launch of the hot path:  ncu --metrics gpu__time_duration.sum ... python scripts/profile_step.py [steps] [gemm mode] [transformer-base | s2s-deep-gru]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = graft.load_package()
lib = pkg.load()
model = sys.argv[3] if len(sys.argv) > 3 else "transformer-base"
if model == "s2s-deep-gru":  # BASELINE.json configs[2]
    opts = {"type": "s2s", "dim-vocabs": [32000, 32000], "dim-emb": 512, "dim-rnn": 1024, "enc-depth": 4, "dec-depth": 4, "enc-cell": "gru", "dec-cell": "gru",
            "cost-type": "ce-mean", "label-smoothing": 0, "optimizer": "adam", "learn-rate": 0.0001, "clip-norm": 1, "seed": 1234, "workspace": 16384, "gemm-mode": mode}
else:
    opts = pkg.transformer_base_options(gemm_mode=mode)
t = lib.trainer(dict(opts, **{"graph-replay": "false"}))
for s in range(steps):
    t.next_synthetic_batch(64, 50, 50)
    t.compute_gradients()
    t.update()
    print("step", s, "cost", t.cost(), flush=True)
print(t.stats())
