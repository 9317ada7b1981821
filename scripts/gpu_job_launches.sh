# GPU job: ncu launch list (gpu__time_duration) of a short S1 bench run -> gpurun_out/launches_<tag>.csv
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TAG=${1:-s1}; shift
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-ref-cuda "$@" > gpurun_out/launches_$TAG.log 2>&1
python - "$TAG" <<'PY'
import csv, sys, collections
rows = list(csv.reader(open(f"gpurun_out/launches_{sys.argv[1]}.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mn, mv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    d = agg.setdefault(r[kn][:70], collections.defaultdict(list))
    d[r[mn]].append(float(r[mv].replace(",", "")))
for k, d in agg.items():
    t = d["gpu__time_duration.sum"]
    print(f"{k:70s} n={len(t):3d} avg_us={sum(t)/len(t)/1e3:8.2f} rd_MB={sum(d['dram__bytes_read.sum'])/len(t)/1e6:8.2f} wr_MB={sum(d['dram__bytes_write.sum'])/len(t)/1e6:8.2f}")
PY
