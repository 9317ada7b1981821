# GPU job: host-buffer step parity tests, then the S1 line (e2e block included)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "update_host" 2>&1 | tail -15
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-ref-cuda --no-extras > gpurun_out/e2e_quick.json 2> gpurun_out/e2e_quick.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/e2e_quick.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"]*1e3,1), d["e2e"])
PY
