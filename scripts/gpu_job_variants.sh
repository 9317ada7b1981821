# GPU job: parity tests (all three collide variants), then S1 with collide tile / PDL on and off, 2^23 with the defaults
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not multigpu" 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for v in "0 0" "0 1" "1 1"; do
  set -- $v
  timeout 300 python bench.py --steps 200 --warmup 50 --no-cpu-baseline --no-ref-cuda --collide-tile $1 --pdl $2 > gpurun_out/bench_s1_t$1_p$2.json 2> gpurun_out/bench_s1_t$1_p$2.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda --robots-log2 23 > gpurun_out/bench_2p23.json 2> gpurun_out/bench_2p23.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_s1_t*.json")) + ["gpurun_out/bench_2p23.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,3), round(d["ms_per_step"]*1e3,1), round(d["back_to_back"]["ms_per_step"]*1e3,1), {k:round(v["avg_us"],1) for k,v in d["stages"].items()}, d["state_finite"])
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json",".err")).read()[-600:])
PY
