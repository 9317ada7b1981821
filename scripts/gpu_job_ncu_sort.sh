# GPU job: ncu --set full of the sort kernels at a given size (default 2^23)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
L=${1:-23}
ncu --set full --import-source on --clock-control none -k regex:"k_onesweep" -s 9 -c 3 -f -o gpurun_out/prof_sort_2p$L \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --robots-log2 $L > gpurun_out/ncu_sort_2p$L.log 2>&1
tail -2 gpurun_out/ncu_sort_2p$L.log | cut -c1-300
