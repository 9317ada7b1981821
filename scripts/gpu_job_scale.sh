# GPU job (N GPUs): the S2 bench line exactly as the driver launches it.  usage: gpu_job_scale.sh N [steps]
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
N=${1:-2}; K=${2:-20}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $K --warmup 5 ) > gpurun_out/bench_scale_${N}x.json 2> gpurun_out/bench_scale_${N}x.err
grep real gpurun_out/bench_scale_${N}x.err
python - "$N" <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/bench_scale_{sys.argv[1]}x.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["config"]["exchange"], {k:round(v["us_per_step"],1) for k,v in d["stages_rank0"].items()}, d["e2e"]["value"], d["clocks"])
PY
