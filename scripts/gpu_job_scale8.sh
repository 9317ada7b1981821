# GPU job (N GPUs): the S2 line exactly as the driver launches it, then the C++ launcher on the same swarm.  usage: gpu_job_scale8.sh N
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
N=${1:-8}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_scale_${N}x.json 2> gpurun_out/bench_scale_${N}x.err
grep real gpurun_out/bench_scale_${N}x.err
python - "$N" <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/bench_scale_{sys.argv[1]}x.json").read().strip().splitlines()[-1])
print(d.get("stages_us_per_rank"))
print(d["n_gpus"], d["value"], d["ms_per_step"], {k:round(v["us_per_step"],1) for k,v in d["stages_rank0"].items()}, "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("cpus_bound_rank0"), d.get("parity_check",{}).get("bit_equal"), d.get("speedup_vs_one_gpu"), d["one_gpu_same_workload"]["ms_per_step"])
PY
cd /tmp && timeout 600 $GRAFT_REPO_ROOT/particlerobotsimulations_b200/ParticleBot $GRAFT_REPO_ROOT/examples/synthetic_s2.cfg --gpus $N --steps 400 --no-csv --quiet 2>&1 | tail -3
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|^CPU(s)" | head
