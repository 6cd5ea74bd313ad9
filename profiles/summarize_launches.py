"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py.

usage: python profiles/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/...
Steps are delimited by the stem kernel (first launch of every encode).  Times under ncu are
cold-cache and serialised: compare SHARES of the step, not absolutes (B200_PROFILING.md).
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("vqae::<unnamed>::", "").replace("void ", "")
    return name.strip()


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append((short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"])))
    starts = [i for i, r in enumerate(rows) if r[0].startswith("stem_in")]
    lo = starts[step]
    hi = starts[step + 1] if step + 1 < len(starts) else len(rows)
    sel = rows[lo:hi]
    total = sum(r[3] for r in sel)
    agg = OrderedDict()
    for name, grid, block, ns in sel:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    print(f"# source: {path}; step {step} of the run = launches [{lo}, {hi}) ({len(sel)} launches)")
    print(f"# total serialised device time {total / 1e6:.3f} ms")
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg_us':>9s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:70s} {n:8d} {ns / 1e6:10.3f} {ns / total:7.1%} {ns / n / 1e3:9.1f}")


if __name__ == "__main__":
    main()
