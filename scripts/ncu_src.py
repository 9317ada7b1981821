#!/usr/bin/env python
"""Summarise an ncu report: per-kernel headline metrics, and (with --src KERNEL) the hottest SASS lines
by stall samples.  Usage: ncu_src.py report.ncu-rep [--src substring] [--top N] [--launch I]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else None
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print(" | ".join(f"{hdr[i].split('.')[0][-28:]}={r[i][:60]}" for i in idx))
if src:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name",')
    sel = [b for b in blocks[1:] if src in b.split("\n")[0]]
    li = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
    b = sel[li]
    rr = list(csv.reader(io.StringIO(b)))
    h = rr[1]
    c = {k: h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Avg. Threads Executed")}
    st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    data = [r for r in rr[2:] if len(r) > c["# Samples"] and r[c["# Samples"]].isdigit()]
    tot = sum(int(r[c["# Samples"]]) for r in data)
    print("kernel:", b.split("\n")[0][:120], "total samples", tot, "instr", sum(int(r[c["Instructions Executed"]]) for r in data))
    order = sorted(range(len(data)), key=lambda i: -int(data[i][c["# Samples"]]))[:top]
    for i in sorted(order):
        r = data[i]
        reasons = sorted(((int(r[h.index(k)] or 0), k[6:]) for k in st), reverse=True)[:3]
        print(f"{i:5d} {r[c['Source']][:60]:60s} smp={r[c['# Samples']]:>6s} exec={r[c['Instructions Executed']]:>8s} thr={r[c['Avg. Threads Executed']]:>3s}", [f"{k}:{v}" for v, k in reasons if v])
