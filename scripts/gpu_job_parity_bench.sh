# GPU job: parity tests, then the S1 bench line and a 2^23 line (no CPU baseline) — used while iterating on kernels
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "not multigpu" 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 30 --no-cpu-baseline > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --robots-log2 23 > gpurun_out/bench_2p23.json 2> gpurun_out/bench_2p23.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_s1.json", "gpurun_out/bench_2p23.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline_step"]["frac"], {k:round(v["avg_us"],1) for k,v in d["stages"].items()})
PY
