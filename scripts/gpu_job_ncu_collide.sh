# GPU job: ncu --set full of ONE collide launch at the evolved state of S1 (after 260 steps)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide_exact --launch-skip 260 --launch-count 1 -f -o gpurun_out/prof_collide_evolved python bench.py --steps 20 --warmup 250 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_collide_evolved.log 2>&1
tail -5 gpurun_out/ncu_collide_evolved.log
ls -la gpurun_out/*.ncu-rep
