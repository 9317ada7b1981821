#!/usr/bin/env python
"""Builds profiles/<tag>_ncu_summary.md, profiles/<tag>_launches_s1.csv and profiles/traffic.json from the ncu
reports and launch list that scripts/gpu_job_profiles.sh left in gpurun_out/ (run here, no GPU needed).
The round tag comes from the environment (ROUND, default r2)."""
import collections, csv, io, json, os, shutil, subprocess
TAG = os.environ.get("ROUND", "r2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
shutil.copy(os.path.join(G, f"launches_{TAG}_s1.csv"), os.path.join(P, f"{TAG}_launches_s1.csv"))
if os.path.exists(os.path.join(G, f"sort_timeline_{TAG}.txt")):
    shutil.copy(os.path.join(G, f"sort_timeline_{TAG}.txt"), os.path.join(P, f"{TAG}_sort_timeline.txt"))

rows = list(csv.reader(open(os.path.join(G, f"launches_{TAG}_s1.csv"))))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mn, mv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > mv:
        agg.setdefault(r[kn], collections.defaultdict(list))[r[mn]].append(float(r[mv].replace(",", "")))
binned = ["k_control_integrate_hash_x2", "k_control_integrate_hash<1, 1>", "k_cell_tile_sums", "k_cell_scan_tiles", "k_cell_apply", "k_cell_scatter",
          "k_reorder_binned", "k_collide_exact<0, 0, prs::PackedLayout, 1>", "k_collide_exact<0, 0, PackedLayout, 1>"]
onesweep = ["k_control_integrate_hash<1, 0>", "k_histogram", "k_onesweep", "k_reorder_packed", "k_max_population",
            "k_collide_exact<0, 0, prs::PackedLayout, 0>", "k_collide_exact<0, 0, PackedLayout, 0>"]
def short(k): return k.split("(")[0].replace("void ", "").replace("prs_bin::", "").replace("prs_sort::", "").replace("prs::", "")
def table(names, passes):
    out, tot = [], 0.0
    for nm in names:
        for k, d in agg.items():
            if nm in k:
                t = d["gpu__time_duration.sum"]; avg = sum(t) / len(t) / 1e3
                mult = passes if "k_onesweep" in nm else 1
                tot += avg * mult
                out.append((short(k), len(t), avg, avg * mult, sum(d["dram__bytes_read.sum"]) / len(t) / 1e6, sum(d["dram__bytes_write.sum"]) / len(t) / 1e6))
    return out, tot

def raw(rep):
    txt = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    return r[0], r[1], r[2:]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
M = [("gpu__time_duration.sum", "time us"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("smsp__inst_executed.sum", "warp inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA %"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
     ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
     ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
def full_table(rep):
    hd, units, rr = raw(rep)
    seen, lines = set(), []
    for r in rr:
        name = short(r[hd.index("Kernel Name")])
        if name in seen: continue
        seen.add(name)
        vals = []
        for m, _ in M:
            v = r[hd.index(m)] if m in hd else ""
            try:
                x = float(v.replace(',', ''))
                if m.startswith("dram__bytes"): x *= UNIT.get(units[hd.index(m)], 1.0) / 1e6   # -> MB whatever unit ncu chose
                v = f"{x:.4g}"
            except ValueError: pass
            vals.append(v)
        # achieved DRAM bandwidth of this launch against the measured peak (MEASURED_PEAKS.json)
        try:
            t_us = float(vals[0]); mb = float(vals[[m for m, _ in M].index("dram__bytes_read.sum")]) + float(vals[[m for m, _ in M].index("dram__bytes_write.sum")])
            gbs = mb / t_us * 1e3
            vals += [f"{gbs:.0f}", f"{100 * gbs / PEAK:.1f}"]
        except (ValueError, ZeroDivisionError):
            vals += ["", ""]
        lines.append((name, vals))
    return lines, hd, units, rr

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
STAGE_NOTE = ""
bj = os.path.join(G, f"bench_{TAG}.json")
if os.path.exists(bj):
    try:
        d = json.loads(open(bj).read().strip().splitlines()[-1])
        st = d["stages"]
        STAGE_NOTE = (f"bench.py stage times of the same build (CUDA events, L2 flushed, timed from step {d.get('timed_state_step')}): "
                      + ", ".join(f"{k} {v['avg_us']:.1f} µs" for k, v in st.items())
                      + f"; whole step with programmatic dependent launch {1e3 * d['ms_per_step']:.1f} µs (under ncu the small kernels pay "
                        "relatively more for cold caches and serialisation).")
    except Exception:
        pass
out = [f"# Round {TAG[1:]} — ncu evidence (B200, S1 = 2^20 robots, sort every step)", "",
       "Generated by `scripts/make_profiles.py` from `scripts/gpu_job_profiles.sh` (reports under `gpurun_out/`).", "",
       "## Launch list of `python bench.py --steps 6 --warmup 4` at the timed state (launches of the first 260 untimed steps skipped; "
       "`ncu --metrics gpu__time_duration.sum,dram__bytes_* --clock-control none`)", "",
       f"Cold-cache, serialised launches: compare SHARES with the CUDA-event stage times of `bench.py`, not absolutes. Raw list: `{TAG}_launches_s1.csv`.", ""]
for title, names, passes in (("binned route (the route of every timed step once the swarm is known to be sparse)", binned, 1),
                             ("onesweep route (first sort steps, crowded swarms, C-ABI sortParticlebots)", onesweep, 3)):
    t, tot = table(names, passes)
    if not t or not any(n.startswith(("k_cell", "k_onesweep", "k_histogram")) for n, *_ in t):
        continue      # this route does not occur in the captured launches
    out += [f"### {title}", "", "| kernel | launches | avg µs | per step µs | share | dram rd MB | dram wr MB |", "|---|---|---|---|---|---|---|"]
    out += [f"| {n} | {c} | {a:.1f} | {s:.1f} | {100 * s / tot:.1f} % | {rd:.1f} | {wr:.1f} |" for n, c, a, s, rd, wr in t]
    out += [f"| **sum** | | | **{tot:.1f}** | | | |", ""]
out += [STAGE_NOTE, ""]
for title, rep in ((f"`ncu --set full` of the step's kernels (prof_{TAG}_step)", f"prof_{TAG}_step.ncu-rep"),
                   (f"`ncu --set full` of the onesweep route's kernels (prof_{TAG}_onesweep)", f"prof_{TAG}_onesweep.ncu-rep")):
    if not os.path.exists(os.path.join(G, rep)):
        continue
    lines, hd, units, rr = full_table(rep)
    out += [f"## {title}", "", "| kernel | " + " | ".join(n for _, n in M) + " | DRAM GB/s | % of measured HBM peak |", "|---|" + "---|" * (len(M) + 2)]
    out += [f"| {n} | " + " | ".join(v) + " |" for n, v in lines]
    out.append("")
    if "step" in rep:
        traffic = {}
        for r in rr:
            name = short(r[hd.index("Kernel Name")])
            if "k_collide" in name:
                rd = float(r[hd.index("dram__bytes_read.sum")].replace(",", "")) * UNIT.get(units[hd.index("dram__bytes_read.sum")], 1.0)
                wr = float(r[hd.index("dram__bytes_write.sum")].replace(",", "")) * UNIT.get(units[hd.index("dram__bytes_write.sum")], 1.0)
                traffic = {"collide": rd + wr, "_what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes) of k_collide_exact at S1, "
                                                        f"ncu --set full (profiles/{TAG}_ncu_summary.md)", "_read": rd, "_write": wr}
                break
        json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
out += ["dram bytes are per launch; at 2^20 robots the whole state (~100 MB) nearly fits the 126 MB L2, so under ncu's serialised replay DRAM traffic "
        "under-reads the algorithmic bytes (collide: 43 B x 2^20 = 45 MB algorithmic vs 33 MB read; its scattered 16 MB of results stay in L2).",
        "collide: XU (MUFU) and FMA pipes ~45-50 % busy each, issue slots ~70 % — a balanced FP32/MUFU bound, DRAM at 2 %.",
        "`r1_collide_evolved.md`: the round-1 capture of the same kernel at step 260 with the per-region instruction and stall-sample shares; "
        "`r2_collide_patch.md`: the pair-sharing patch kernel against it.", ""]
open(os.path.join(P, f"{TAG}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
