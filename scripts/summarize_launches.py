"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel
table (launches per step, total / mean device time, share of the step).

  python scripts/summarize_launches.py gpurun_out/launches.csv [steps] > profiles/launches_rNN.md
Per-launch times under ncu are cold-cache and serialised: read the SHARES, not the absolutes.
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"marian::", "", name)
    name = re.sub(r"functional::", "", name)
    name = re.sub(r"^void ", "", name)
    base = name.split("(")[0]
    return base if len(base) <= 120 else base[:117] + "..."


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    agg = defaultdict(lambda: [0, 0.0])
    n = 0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r.get("Metric Unit", "ns"), 1)
        a = agg[short(r["Kernel Name"])]
        a[0] += 1
        a[1] += ns
        n += 1
    total = sum(v[1] for v in agg.values())
    print("| kernel | launches/step | us/step | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for name, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %.1f | %.1f | %.1f | %.1f%% |" % (name, c / steps, ns / steps / 1e3, ns / c / 1e3, 100 * ns / total))
    print("\ntotal: %.1f launches/step, %.3f ms/step device time (sum of serialised, cold-cache launches; %d steps profiled)" % (n / steps, total / steps / 1e6, steps))


if __name__ == "__main__":
    main()
