# GPU job (2 GPUs): multi-rank parity tests (torch.distributed slabs and the C++ launcher on real devices), the S2 line as the
# driver launches it, and the C++ launcher on the same swarm (ParticleBot --gpus 2)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -4
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/bench_scale_2x.json 2> gpurun_out/bench_scale_2x.err
grep real gpurun_out/bench_scale_2x.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_scale_2x.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["config"]["exchange"], {k:round(v["us_per_step"],1) for k,v in d["stages_rank0"].items()}, d["e2e"]["value"], d.get("parity_check"), d.get("speedup_vs_one_gpu"))
PY
cd /tmp && for g in 2; do timeout 600 $GRAFT_REPO_ROOT/particlerobotsimulations_b200/ParticleBot $GRAFT_REPO_ROOT/examples/synthetic_s2.cfg --gpus $g --steps 300 --no-csv --quiet 2>&1 | tail -2; done
