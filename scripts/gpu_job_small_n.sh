# GPU job: kernel launch list of the headless runner on an example cfg (reference cadence), fused backend
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
CFG=${1:-example}
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_$CFG.csv \
  particlerobotsimulations_b200/ParticleBot examples/$CFG.cfg --steps 100 --no-csv --quiet > gpurun_out/launches_$CFG.log 2>&1
python - "$CFG" <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(f"gpurun_out/launches_{sys.argv[1]}.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > mv: agg.setdefault(r[kn][:80], []).append(float(r[mv].replace(",", "")))
for k, t in agg.items(): print(f"{k:80s} n={len(t):3d} avg_us={sum(t)/len(t)/1e3:8.2f}")
PY
particlerobotsimulations_b200/ParticleBot examples/$CFG.cfg --steps 20000 --no-csv --quiet
