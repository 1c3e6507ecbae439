"""Per-parameter gradient error of the CUDA step vs the CPU oracle (debug aid)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load(); orc = graft.load_oracle()
OPTS = {"tr": ("type=transformer;dim-vocabs=200,220;dim-emb=64;transformer-heads=4;transformer-dim-ffn=128;enc-depth=2;dec-depth=2;workspace=256"),
        "gru": "type=s2s;dim-vocabs=200,220;dim-emb=32;dim-rnn=64;enc-depth=2;dec-depth=2;workspace=256"}
which = sys.argv[1] if len(sys.argv) > 1 else "tr"
extra = sys.argv[2] if len(sys.argv) > 2 else ""
def grads(l, mode, extra=""):
    t = l.trainer(OPTS[which] + ";gemm-mode=%d;graph-replay=false%s" % (mode, extra))
    t.next_synthetic_batch(8, 11, 13, padded=True)
    t.compute_gradients(keep_logits=True)
    g = {n: t.get_tensor(n, grad=True) for n, _ in t.param_names()}
    c = t.cost(); t.close(); return c, g
ce, ge = grads(orc, 0)
cg, gg = grads(lib, 0, extra)
print("cost", ce, cg)
bad = 0
for n in ge:
    err = float(np.abs(gg[n].astype(np.float64) - ge[n]).max()); sc = float(np.abs(ge[n]).max())
    flag = "BAD" if err > 1e-3 * max(sc, 1e-3) else "ok"
    bad += flag == "BAD"
    print("%-40s err %.3e scale %.3e %s" % (n, err, sc, flag))
print("bad:", bad, "of", len(ge))
