# GPU job: selected parity tests, then the S1 line with the library defaults (no CPU baseline / reference-kernel block)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "${1:-binning or large_swarm}" 2>&1 | tail -5
timeout 300 python bench.py --steps 200 --warmup 50 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_s1_quick.json 2> gpurun_out/bench_s1_quick.err
python - <<'PY'
import json
for f in ["gpurun_out/bench_s1_quick.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,3), round(d["ms_per_step"]*1e3,1), round(d["back_to_back"]["ms_per_step"]*1e3,1), {k:round(v["avg_us"],1) for k,v in d["stages"].items()}, d["state_finite"], d["secondary"]["ms_per_step"])
        print(d["e2e"])
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json",".err")).read()[-600:])
PY
