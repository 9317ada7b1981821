# GPU job (round 2): the ncu evidence for profiles/ at the TIMED state of the bench (step 260): launch list of the step's kernels,
# one --set full capture of each of them, and the bench line of the same build beside them.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
# 260 untimed steps x 6 kernels of the binned route come first (the first 4 steps take the onesweep route: a few more launches)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 1640 -c 300 --csv \
  --log-file gpurun_out/launches_r2_s1.csv python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-ref-cuda --no-extras > gpurun_out/launches_r2_s1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_collide|k_control_integrate|k_cell_|k_reorder_binned" --launch-skip 1640 -c 7 -f \
  -o gpurun_out/prof_r2_step python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-ref-cuda --no-extras > gpurun_out/ncu_r2_step.log 2>&1
tail -2 gpurun_out/ncu_r2_step.log; ls -la gpurun_out/prof_r2_step.ncu-rep gpurun_out/launches_r2_s1.csv
