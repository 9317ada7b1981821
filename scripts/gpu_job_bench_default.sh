# GPU job: the two bench arms exactly as the driver launches them (N=1)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
( time python bench.py --impl reference --gpus 1 ) > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
( time python bench.py --gpus 1 ) > gpurun_out/bench_native_arm.json 2> gpurun_out/bench_native_arm.err
tail -4 gpurun_out/bench_reference_arm.err; cat gpurun_out/bench_reference_arm.json | cut -c1-700
tail -4 gpurun_out/bench_native_arm.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_native_arm.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","roofline","roofline_step","cpu_baseline","ref_cuda","clocks","gpu_launches"): print(k, d[k])
print({k:round(v["avg_us"],1) for k,v in d["stages"].items()})
PY
