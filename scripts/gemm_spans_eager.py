"""Per-launch kernel spans of the GEMM family for an EAGER step (no CUDA graph: kernels run one by one with host gaps)
next to the replayed step - separates cache / HBM effects from in-graph concurrency effects."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
os.environ["MRN_GEMM_SPANS"] = "1"
fn = lib.c.mrn_gemm_profile
fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_size_t)]
for replay in ("false", "true"):
    o = pkg.transformer_base_options(gemm_mode=4)
    o["graph-replay"] = replay
    t = lib.trainer(o)
    def step():
        t.next_synthetic_batch(64, 50, 50); t.compute_gradients(); t.update(); t.cost()
    step(); step()
    ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_size_t()
    os.environ["MRN_GEMM_PROFILE_DUMP"] = "gpurun_out/spans_%s.csv" % ("replay" if replay == "true" else "eager")
    fn(1, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
    for _ in range(3):
        step()
    fn(0, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
    print("graph-replay=%s: %d launches, %.3f ms GEMM spans per step" % (replay, n.value, ms.value), flush=True)
    t.close()
