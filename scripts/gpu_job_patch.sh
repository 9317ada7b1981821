# GPU job: patch-kernel probe (stage times with / without, patch statistics) + one ncu --set full capture of k_collide_patch
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
python scripts/patch_probe.py > gpurun_out/patch_probe.json 2> gpurun_out/patch_probe.err; cat gpurun_out/patch_probe.json; tail -3 gpurun_out/patch_probe.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collide_patch --launch-skip 262 --launch-count 1 -f -o gpurun_out/prof_patch python scripts/patch_probe.py --only 1 --steps 4 > gpurun_out/ncu_patch.log 2>&1
tail -3 gpurun_out/ncu_patch.log
ls -la gpurun_out/*.ncu-rep
