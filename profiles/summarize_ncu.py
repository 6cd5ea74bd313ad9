"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + stall-sample shares.
usage: python profiles/summarize_ncu.py <report.ncu-rep> "<header line>" > summary.txt"""
import csv
import subprocess
import sys

rep, header = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max")
print("# " + header)
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:85s} {u:16s} {v}")
stall = {h: float(v) for h, v in zip(hdr, vals)
         if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v}
tot = sum(stall.values()) or 1.0
print("# warp-state samples (share of all samples)")
for h, v in sorted(stall.items(), key=lambda kv: -kv[1]):
    if v / tot >= 0.01:
        print(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', ''):30s} {100 * v / tot:5.1f} %")
