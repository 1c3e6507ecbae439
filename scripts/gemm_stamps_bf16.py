"""Timeline of the persistent bf16 GEMM (csrc/kernels/gemm.cu gGemmBf16Persistent): per-CTA %globaltimer stamps of
ONE launch - per tile: first TMA issued, first operands seen by the MMA thread, last MMA issued, accumulator complete,
epilogue done, last TMA issued - summarised over the CTAs (medians, microseconds relative to the kernel start)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
pkg = graft.load_package(); lib = pkg.load()
st = torch.cuda.Stream(); torch.cuda.set_stream(st); lib.set_stream(st.cuda_stream)
g = lib.gemm(4)
rs = np.random.RandomState(0)
cases = [(3200, 512, 2048, 512, 0, 1, 0.0, "ffn gated dH"), (3200, 512, 512, 2048, 0, 0, 0.0, "ffn up fwd"), (3200, 512, 512, 32000, 0, 0, 0.0, "logits fwd"), (3200, 512, 3200, 32000, 1, 0, 1.0, "logits dW")]
for (ra, ca, rb, cb, tA, tB, beta, label) in cases:
    M = ca if tA else ra; N = rb if tB else cb
    A = lib.array(rs.standard_normal((ra, ca)).astype(np.float32)); B = lib.array(rs.standard_normal((rb, cb)).astype(np.float32)); C = lib.zeros((M, N))
    gated = "gated" in label
    H = lib.array(rs.standard_normal((M, N)).astype(np.float32)) if gated else None
    def run():
        if gated:
            lib.call("mrn_prod_swish_grad_nt", g.h, C.t(), A.t(), B.t(), H.t(), beta)
        else:
            lib.call("mrn_prod", g.h, C.t(), A.t(), B.t(), tA, tB, beta, 1.0)
    stamps = torch.zeros(64 * 148, dtype=torch.int64, device="cuda")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    lib.call("mrn_gemm_debug_stamps", stamps.data_ptr())
    run()
    lib.call("mrn_gemm_debug_stamps", None)
    torch.cuda.synchronize()
    s = stamps.cpu().numpy().reshape(-1, 64).astype(np.float64)
    s = s[s[:, 0] > 0]
    t0 = s[:, 0].min()
    rel = np.where(s > 0, (s - t0) / 1000.0, np.nan)
    out = {"case": label, "ctas": int(len(s)), "kernel_span_us": float(np.nanmax(rel[:, 63])), "start_p50": float(np.nanmedian(rel[:, 0])),
           "prologue_done_p50": float(np.nanmedian(rel[:, 1]))}
    tiles = []
    for j in range(10):
        b = 2 + 6 * j
        if np.all(np.isnan(rel[:, b])):
            break
        tiles.append({k: round(float(np.nanmedian(rel[:, b + o])), 2) for k, o in (("tma_first", 0), ("mma_first_operands", 1), ("mma_last_issued", 2), ("acc_complete", 3), ("epilogue_done", 4), ("tma_last", 5))})
    out["tiles_p50"] = tiles
    print(json.dumps(out), flush=True)
