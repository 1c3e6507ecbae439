import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
which = sys.argv[1]
if which == "tiny":
    o = "type=transformer;dim-vocabs=200,224;dim-emb=64;transformer-heads=4;transformer-dim-ffn=128;enc-depth=2;dec-depth=2;workspace=256;gemm-mode=4;graph-replay=%s" % sys.argv[2]
    shape = (8, 11, 13)
else:
    o = pkg.transformer_base_options(gemm_mode=4); o["graph-replay"] = sys.argv[2]
    shape = (64, 50, 50)
t = lib.trainer(o)
for s in range(int(sys.argv[3])):
    t.next_synthetic_batch(*shape, padded=True)
    t.compute_gradients(); t.update()
    print(s, t.batch_words(), t.cost(), flush=True)
print("ok")
