# GPU job: headless runner on the example cfgs, 20000 steps, fused backend (reference cadence)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for CFG in example example_dead_cells example_obstacle example_gap example_object_transport; do
  echo "== $CFG fused"
  particlerobotsimulations_b200/ParticleBot examples/$CFG.cfg --steps 20000 --no-csv --quiet
done 2>&1 | tee gpurun_out/small_n_quick.log
