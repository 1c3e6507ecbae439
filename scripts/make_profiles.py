"""Turns the raw artefacts a GPU run left in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/make_profiles.py r01
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(PROF, exist_ok=True)


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def write(name, text):
    with open(os.path.join(PROF, name), "w") as fh:
        fh.write(text)
    print("wrote profiles/" + name)


# 1. launch list of two eager steps
if os.path.exists(os.path.join(OUT, "launches.csv")):
    body = run([sys.executable, os.path.join(ROOT, "scripts", "summarize_launches.py"), os.path.join(OUT, "launches.csv"), "2"])
    write("launches_%s.md" % tag, "# Kernel launch list of the Transformer-base training step (tf32 mode, eager, 2 steps)\n\n"
          "`ncu --metrics gpu__time_duration.sum --clock-control none python scripts/profile_step.py 2 3` (scripts/gpu_bench.sh, NCU=1).\n"
          "Per-launch times under ncu are serialised and cold-cache: read the SHARES.  The replayed step itself is timed by bench.py.\n\n" + body)

# 2. in-graph GEMM table
p = os.path.join(OUT, "gemm_dump_side.csv.spans")
if os.path.exists(p):
    body = run([sys.executable, os.path.join(ROOT, "scripts", "summarize_gemm_dump.py"), p])
    write("gemm_kernel_spans_%s.md" % tag, "# Tensor-core GEMM launches of ONE replayed step: kernel execution spans\n\n"
          "`MRN_GEMM_PROFILE_DUMP=gpurun_out/gemm_dump_side.csv python bench.py` (second profiling pass, MRN_GEMM_SPANS): every CTA of a launch folds\n"
          "its %globaltimer start / end into a per-launch min / max slot (csrc/kernels/gemm.cu), so the duration is first-CTA-start ..\n"
          "last-CTA-end INSIDE the replayed graph, with the side stream running concurrently and no extra graph nodes.  This is the\n"
          "duration bench.py's `roofline` uses.  Weight-gradient products (op(A)=T) run on the side stream.\n\n" + body)
for f in ("gemm_dump_side.csv",):
    p = os.path.join(OUT, f)
    if os.path.exists(p):
        body = run([sys.executable, os.path.join(ROOT, "scripts", "summarize_gemm_dump.py"), p])
        write("gemm_in_graph_%s.md" % tag, "# Tensor-core GEMM launches of ONE replayed step, timed inside the CUDA graph\n\n"
              "`MRN_GEMM_PROFILE_DUMP=gpurun_out/gemm_dump_side.csv python bench.py` - CUDA events recorded as external event nodes around\n"
              "every launch of the captured step (csrc/kernels/gemm.cu ProfileScope).  Each pair includes ~6.7 us of event-record node\n"
              "latency (profiles/graph_gap_probe_r01.json): compare with gemm_kernel_spans_rNN.md.\n"
              "Weight-gradient products (op(A)=T) run on the side stream, concurrently with the main chain.\n\n" + body)

# 3. DRAM traffic of the GEMM family
p = os.path.join(OUT, "gemm_dram.csv")
if os.path.exists(p):
    lines = [ln for ln in open(p) if not ln.startswith("==")]
    per = {}
    for r in csv.DictReader(lines):
        k = (r["ID"], r["Kernel Name"].split("(")[0][-40:], r.get("Grid Size", ""))
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
        per.setdefault(k, {})[r["Metric Name"]] = v * scale
    n = len(per)
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values())
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values())
    ns = sum(d.get("gpu__time_duration.sum", 0) for d in per.values())
    steps = 2
    text = ("# DRAM traffic of the tensor-core GEMM family (gGemmTf32), one eager step\n\n"
            "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:gGemm python scripts/profile_step.py 2 3`\n\n"
            "| launches/step | DRAM read MB/step | DRAM write MB/step | device ms/step (serialised) |\n|---:|---:|---:|---:|\n"
            "| %.0f | %.1f | %.1f | %.3f |\n\n" % (n / steps, rd / steps / 1e6, wr / steps / 1e6, ns / steps / 1e6))
    text += ("Algorithmic bytes of the same launches (read A and B once, write C once, +read C when accumulating), fp32: see DESIGN.md 4.\n"
             "Writes mostly stay in the 126 MB L2 at kernel end (write-back happens later), hence the small DRAM write figure.\n")
    write("gemm_dram_%s.md" % tag, text)
    json.dump({"launches_per_step": n / steps, "dram_read_bytes_per_step": rd / steps, "dram_write_bytes_per_step": wr / steps},
              open(os.path.join(PROF, "gemm_dram_%s.json" % tag), "w"))

# 4. ncu --set full captures: selected raw metrics
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for f in sorted(os.listdir(OUT)):
    if not f.endswith(".ncu-rep"):
        continue
    raw = run(["ncu", "-i", os.path.join(OUT, f), "--page", "raw", "--csv"])
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if any(h == w or h.endswith("." + w) for w in WANT)]
    name = f[:-len(".ncu-rep")]
    text = "# ncu --set full --clock-control none --import-source on: %s\n\nCapture of launches inside `python scripts/profile_step.py 2 3` (eager Transformer-base step, tf32 mode); scripts/gpu_bench.sh NCUFULL.\n\n" % name
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    for r in rows[2:]:
        text += "## %s grid %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (r[ki].split("(")[0].replace("void ", ""), r[gi])
        for i in cols:
            text += "| %s | %s | %s |\n" % (hdr[i], r[i], units[i])
        text += "\n"
    write("ncu_%s_%s.md" % (name.replace("prof_", ""), tag), text)

# 5. per-operator table
p = os.path.join(OUT, "op_table.json")
if os.path.exists(p):
    rows = [json.loads(ln) for ln in open(p) if ln.strip().startswith("{")]
    rows = [r for r in rows if "op" in r]
    text = ("# Per-operator roofline table at the config-B shapes (this repo's sm_100a kernels next to the reference's own kernels recompiled for sm_100a)\n\n"
            "`python scripts/op_bench.py` (CUDA events, buffers rotate through a pool larger than L2).  HBM peak / bf16 peak from MEASURED_PEAKS.json.\n\n"
            "| operator | ms | GB/s or TFLOP/s | frac of peak | reference kernel ms | speed-up vs reference kernel |\n|---|---:|---:|---:|---:|---:|\n")
    for r in rows:
        rate = r.get("gbs", r.get("tflops", 0.0))
        frac = r.get("frac_of_hbm_peak", r.get("frac_of_bf16_peak", 0.0))
        text += "| %s | %.4f | %.1f | %.3f | %s | %s |\n" % (r["op"], r["ms"], rate, frac, ("%.4f" % r["ref_ms"]) if "ref_ms" in r else "-",
                                                           ("%.2f" % r["speedup_vs_ref_kernel"]) if "speedup_vs_ref_kernel" in r else "-")
    write("op_table_%s.md" % tag, text)

# 6. the bench line itself
p = os.path.join(OUT, "bench.json")
if os.path.exists(p):
    txt = open(p).read().strip().splitlines()
    if txt:
        write("bench_%s.json" % tag, txt[-1] + "\n")
