# GPU job: selected gpu tests (-k expression in $1), tail of the output
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -25
