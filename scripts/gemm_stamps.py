"""Where does a small tf32 GEMM spend its time?  Per-CTA %globaltimer stamps (start, prologue done,
first operands landed, accumulator complete, end) of ONE launch, summarised over the CTAs."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
st = torch.cuda.Stream(); torch.cuda.set_stream(st); lib.set_stream(st.cuda_stream)
g = lib.gemm(3)
rs = np.random.RandomState(0)
for (ra, ca, rb, cb, tA, tB, beta, label) in [(3200, 512, 512, 512, 0, 0, 0.0, "proj fwd"), (3200, 512, 512, 512, 0, 1, 1.0, "proj dX"),
                                              (3200, 512, 512, 2048, 0, 0, 0.0, "ffn1 fwd"), (3200, 2048, 2048, 512, 0, 0, 0.0, "ffn2 fwd")]:
    M = ra; N = rb if tB else cb
    A = lib.array(rs.standard_normal((ra, ca)).astype(np.float32)); B = lib.array(rs.standard_normal((rb, cb)).astype(np.float32)); C = lib.zeros((M, N))
    stamps = torch.zeros(5 * 4096, dtype=torch.int64, device="cuda")
    for _ in range(3):
        lib.call("mrn_prod", g.h, C.t(), A.t(), B.t(), tA, tB, beta, 1.0)
    torch.cuda.synchronize()
    lib.call("mrn_gemm_debug_stamps", stamps.data_ptr())
    lib.call("mrn_prod", g.h, C.t(), A.t(), B.t(), tA, tB, beta, 1.0)
    lib.call("mrn_gemm_debug_stamps", None)
    torch.cuda.synchronize()
    s = stamps.cpu().numpy().reshape(-1, 5)
    s = s[s[:, 0] > 0]
    t0 = s[:, 0].min()
    rel = (s - t0) / 1000.0
    seg = np.diff(s, axis=1) / 1000.0
    print(json.dumps({"case": label, "ctas": int(len(s)), "kernel_span_us": float(rel[:, 4].max()),
                      "cta_start_us_p50_max": [float(np.median(rel[:, 0])), float(rel[:, 0].max())],
                      "prologue_us_p50_max": [float(np.median(seg[:, 0])), float(seg[:, 0].max())],
                      "first_operands_us_p50_max": [float(np.median(seg[:, 1])), float(seg[:, 1].max())],
                      "mainloop_us_p50_max": [float(np.median(seg[:, 2])), float(seg[:, 2].max())],
                      "epilogue_us_p50_max": [float(np.median(seg[:, 3])), float(seg[:, 3].max())]}))
