# GPU job: the ncu evidence for profiles/: launch list of a short S1 bench run, one --set full capture of the step's
# kernels (binned route; FULL=1 adds the onesweep route and the sort's phase timeline, unchanged since they were taken)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
bash scripts/gpu_job_launches.sh r1_s1 > gpurun_out/launches_r1_s1.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_collide|k_control_integrate|k_cell_|k_reorder_binned" -s 40 -c 7 -f -o gpurun_out/prof_r1_step \
  python bench.py --steps 6 --warmup 12 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_r1_step.log 2>&1
if [ -n "$FULL" ]; then
ncu --set full --clock-control none -k regex:"k_onesweep|k_histogram|k_reorder_packed" -s 4 -c 5 -f -o gpurun_out/prof_r1_onesweep \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_r1_onesweep.log 2>&1
(python scripts/sort_timeline.py 20 22 0; python scripts/sort_timeline.py 23 24 0; python scripts/sort_timeline.py 26 26 0) 2>&1 | grep -v deciles > gpurun_out/sort_timeline_r1.txt
fi
tail -16 gpurun_out/launches_r1_s1.txt
