# GPU job: what the driver runs at round end on one GPU — smoke, the gpu test suite, both bench arms
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python -m pytest tests -x -q -m gpu ) 2>&1 | tail -6
( time python bench.py --impl reference --gpus 1 ) > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err
( time python bench.py --gpus 1 ) > gpurun_out/bench_native_arm.json 2> gpurun_out/bench_native_arm.err
grep real gpurun_out/bench_reference_arm.err gpurun_out/bench_native_arm.err
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_reference_arm.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/bench_native_arm.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"])
for k in ("value","ms_per_step","e2e","roofline","roofline_step","cpu_baseline","ref_cuda","clocks","gpu_launches"): print(k, d[k])
print({k:round(v["avg_us"],1) for k,v in d["stages"].items()})
print("e2e speed-up vs reference arm:", d["e2e"]["value"]/r["value"], " device-resident:", d["value"]/r["value"])
PY
