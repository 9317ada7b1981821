# GPU job: example cfgs through the three backends (headless runner), then single-GPU bench lines at 2^23 and 2^26
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
B=particlerobotsimulations_b200/ParticleBot
rm -f gpurun_out/small_n.log
for c in example example_dead_cells example_obstacle example_gap example_object_transport; do
  for be in fused percall ext:oracle/_ref/libprs_refcuda.so; do
    echo "== $c $be" >> gpurun_out/small_n.log
    $B examples/$c.cfg --steps 20000 --no-csv --quiet --backend $be 2>> gpurun_out/small_n.log
  done
done
python bench.py --robots-log2 23 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_2p23.json 2> gpurun_out/bench_2p23.err
python bench.py --robots-log2 26 --steps 8 --warmup 4 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_2p26.json 2> gpurun_out/bench_2p26.err
python scripts/cpu_oracle_cfgs.py 2>/dev/null | tail -5 >> gpurun_out/small_n.log
grep -A1 "fused\|ext:" gpurun_out/small_n.log | grep ParticleBot | head -12
python - <<'PY'
import json
for f in ("gpurun_out/bench_2p23.json", "gpurun_out/bench_2p26.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["roofline_step"]["frac"], {k:round(v["avg_us"],1) for k,v in d["stages"].items()}, d["e2e"]["value"])
PY
