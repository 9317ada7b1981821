"""Build libparticlebot_b200.so and the headless ParticleBot runner in-tree with nvcc for sm_100a.

Explicit nvcc commands (no torch.utils.cpp_extension, no JIT cache): the built files live next to
the sources, are git-ignored, and travel to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libparticlebot_b200.so")
RUNNER = os.path.join(HERE, "ParticleBot")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]

LIB_SOURCES = ["prs_kernels.cu", "prs_config.cpp", "prs_particlebot.cpp", "prs_video.cpp", "prs_multi.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    out = []
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            out.append(os.path.join(d, f))
    return out


def build(force=False, verbose=False):
    deps = _deps()
    if force or _newer(LIB, deps):
        cmd = [NVCC] + ARCH + COMMON + ["-Xptxas", "-v"] * bool(verbose) + [
            "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xlinker", "-Bsymbolic", "-o", LIB,
        ] + [os.path.join(CSRC, s) for s in LIB_SOURCES] + ["-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    if force or _newer(RUNNER, deps + [LIB]):
        cmd = [NVCC] + ARCH + COMMON + ["-o", RUNNER, os.path.join(CSRC, "prs_main.cpp"), LIB,
                                         "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
