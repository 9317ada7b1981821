cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
(python scripts/sort_timeline.py 20 22 512; python scripts/sort_timeline.py 23 24 512; python scripts/sort_timeline.py 26 26 512) 2>&1 | grep -v "deciles" | tee gpurun_out/sort_timeline.log
