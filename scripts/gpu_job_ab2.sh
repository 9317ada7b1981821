# GPU job: like gpu_job_ab.sh, with the secondary blocks (reference cadence = steps without a sort) in the printout
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "${PRS_AB_TESTS:-collide or trajectory or large_swarm or binning or s1_}" 2>&1 | tail -3
for v in default "$@"; do
  lib=""; [ "$v" != default ] && lib="$PWD/particlerobotsimulations_b200/variants/libparticlebot_b200_$v.so"
  PRS_LIB=$lib timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-ref-cuda --no-s2 > gpurun_out/ab_${v}.json 2> gpurun_out/ab_${v}.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"]*1e3,1), {k:round(v["avg_us"],1) for k,v in d["stages"].items()}, "ref cadence ms", round(d["secondary"]["ms_per_step"],4), "pitch .155", d.get("secondary_spec_pitch_0155",{}).get("ms_per_step"))
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json",".err")).read()[-600:])
PY
