# GPU job (N GPUs): multi-rank parity test, then the slab bench at 2^L robots.  usage: gpu_job_multi.sh N L [steps]
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
N=${1:-2}; L=${2:-24}; K=${3:-20}
python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $K --warmup 5 --robots-log2 $L > gpurun_out/bench_multi_${N}x_2p$L.json 2> gpurun_out/bench_multi_${N}x_2p$L.err
tail -1 gpurun_out/bench_multi_${N}x_2p$L.json | cut -c1-400; tail -1 gpurun_out/bench_multi_${N}x_2p$L.json | grep -o '"slabs".*'; tail -3 gpurun_out/bench_multi_${N}x_2p$L.err
