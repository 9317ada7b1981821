# GPU job: one ncu --set full capture of the collide kernel (S1 bench shape), report into gpurun_out/
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:k_collide -s 5 -c 2 -f -o gpurun_out/prof_collide \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_collide.log 2>&1
tail -3 gpurun_out/ncu_collide.log
