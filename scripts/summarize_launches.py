"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel
table (launch count, total/mean device time, share of the step).

  python scripts/summarize_launches.py gpurun_out/launches.csv [first_step_marker] > profiles/launches_rNN.md
Per-launch times under ncu are cold-cache and serialised: read the SHARES, not the absolutes.
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"marian::", "", name)
    m = re.match(r"(?:void )?([\w:]+)", name)
    base = m.group(1) if m else name
    if "gElementwise" in name:
        acc = "ACC" if re.search(r"gElementwise<\(int\)\d+, \(bool\)1", name) else "SET"
        k = re.search(r"gElementwise<\(int\)(\d+)", name)
        return "ew::gElementwise<K=%s,%s>" % (k.group(1) if k else "?", acc)
    if "gGemmTcgen05" in name:
        bn = re.search(r"gGemmTcgen05<\(int\)(\d+)", name)
        return "gGemmTcgen05<BN=%s>" % (bn.group(1) if bn else "?")
    for key in ("gAddGeneric", "gAddReduceRows"):
        if key in name:
            return "ew::" + key
    return base


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        a = agg[short(name)]
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f | %.1f%% |" % (name, n, ns / 1e6, ns / n / 1e3, 100 * ns / total))
    print("\ntotal: %d launches, %.3f ms device time (sum of serialised, cold-cache launches)" % (len(rows), total / 1e6))


if __name__ == "__main__":
    main()
