"""Runs a few EAGER (no graph replay) Transformer-base training steps so that ncu can list
every kernel launch of the hot path:  ncu --metrics gpu__time_duration.sum ... python scripts/profile_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pkg = graft.load_package()
lib = pkg.load()
t = lib.trainer(dict(pkg.transformer_base_options(gemm_mode=mode), **{"graph-replay": "false"}))
for s in range(steps):
    t.next_synthetic_batch(64, 50, 50)
    t.compute_gradients()
    t.update()
    print("step", s, "cost", t.cost(), flush=True)
print(t.stats())
