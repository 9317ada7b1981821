# GPU job: sort parity tests and bench.py --sort-only with the default build and the named variants (scripts/build_variant.sh)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for v in default "$@"; do
  lib=""; [ "$v" != default ] && lib="$PWD/particlerobotsimulations_b200/variants/libparticlebot_b200_$v.so"
  echo "== $v"
  PRS_LIB=$lib timeout 600 python -m pytest tests -m gpu -x -q -k "sort or binning_route or tiny_swarms" 2>&1 | tail -2
  PRS_LIB=$lib timeout 600 python bench.py --sort-only > gpurun_out/sort_$v.json 2> gpurun_out/sort_$v.err
  python - "$v" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/sort_{sys.argv[1]}.json").read().strip().splitlines()[-1])
for s in d["sizes"] if "sizes" in d else d.get("sort_only", d).get("sizes", []):
    print(s["pairs"], s["key_bits"], "onesweep ms", round(s["onesweep"]["ms"], 4), "cub ms", round(s["cub_key_bits"]["ms"], 4), "ratio", round(s["onesweep_vs_cub_key_bits"], 3), s["identical_output"])
PY
done
