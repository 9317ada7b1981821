cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
B=particlerobotsimulations_b200/ParticleBot
for c in example example_dead_cells example_obstacle example_gap example_object_transport; do
  for be in fused percall ext:oracle/_ref/libprs_refcuda.so; do
    echo "== $c $be" >> gpurun_out/small_n.log
    $B examples/$c.cfg --steps 20000 --no-csv --quiet --backend $be 2>> gpurun_out/small_n.log
  done
done
python bench.py --robots-log2 23 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_2p23.json 2> gpurun_out/bench_2p23.err
python bench.py --robots-log2 26 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2p26.json 2> gpurun_out/bench_2p26.err
tail -3 gpurun_out/small_n.log; cat gpurun_out/bench_2p23.json | cut -c1-600
