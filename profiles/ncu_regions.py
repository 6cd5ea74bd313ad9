"""Summarise an ncu report's source page into contiguous SASS regions with equal execution counts
(= loop bodies): instructions executed, share, stall samples, opcode mix; plus headline metrics.
usage: python profiles/ncu_regions.py report.ncu-rep"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "launch__grid_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:75s} {v:>16s} {u}")
stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(v) for h, v in zip(hdr, vals)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued")}
tot = sum(stalls.values()) or 1
print("stalls:", ", ".join(f"{k} {v / tot:.0%}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
data = rows[2:]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
total = sum(int(r[ia]) for r in data)
segs, prev, start, acc = [], None, 0, 0
for k, r in enumerate(data):
    c = int(r[ia])
    if prev is None or abs(c - prev) > 0.02 * max(c, prev, 1):
        if prev is not None:
            segs.append((start, k - 1, prev, acc))
        start, acc = k, 0
    acc += c
    prev = c
segs.append((start, len(data) - 1, prev, acc))
print(f"total warp instructions {total}")
for s, e, c, a in segs:
    if a > 0.01 * total:
        ops = {}
        for r in data[s:e + 1]:
            toks = r[isrc].split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:9]
        samp = sum(int(r[isamp]) for r in data[s:e + 1])
        print(f"sass {s:4d}-{e:4d} n={e - s + 1:4d} exec/instr={c:9d} share={a / total:6.1%} samples={samp:6d} {top}")
