#!/bin/bash
# Build a compile-time variant of the product library for A/B timing: scripts/build_variant.sh <name> <nvcc -D flags...>
# -> particlerobotsimulations_b200/variants/libparticlebot_b200_<name>.so (git-ignored; select with PRS_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p particlerobotsimulations_b200/variants
C=particlerobotsimulations_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Iinclude -I$C "$@" \
  -shared -Xcompiler -fPIC,-fvisibility=hidden -Xlinker -Bsymbolic \
  -o particlerobotsimulations_b200/variants/libparticlebot_b200_$name.so $C/prs_kernels.cu $C/prs_config.cpp $C/prs_particlebot.cpp $C/prs_video.cpp $C/prs_multi.cpp -ldl -lpthread
echo particlerobotsimulations_b200/variants/libparticlebot_b200_$name.so
