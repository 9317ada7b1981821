"""Generates tests/golden/*.npz by running the REFERENCE's own kernels (oracle/_ref/libprs_refcuda.so:
particlebot_cuda.cu + particlebot_kernel_impl.cuh compiled verbatim for sm_100a) on a B200, driven
by this repo's headless Particlebot host logic in EXTERNAL-backend mode.  Since round 2 the reference's
OWN host class runs headless too (oracle/_ref/libprs_refhost.so: particlebot.cpp compiled verbatim over
a buffer-object stand-in) and tests/test_refhost_gpu.py checks that it reproduces every file here bit for
bit — the goldens are outputs of the reference, whichever of the two host classes drives the kernels.

Run on the GPU box:   python tests/golden/make_golden.py gpurun_out/golden
then copy the files into tests/golden/.  The reference ships no golden vectors of its own
(SURVEY.md §4), so these are the pins for both the CPU oracle and the CUDA path.

Per cfg: snapshots after steps 1, 10, 50, 100 of pos, vel, rad, phase, dead, hash, index and a
digest of the cell tables (occupied cells with their start/end), with the reference cadence
(sort on step 0 only) and with sort_interval = timestep (sort every step).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import particlerobotsimulations_b200 as prs  # noqa: E402
from oracle import binding as ob  # noqa: E402

CFGS = ["example", "example_dead_cells", "example_obstacle", "example_gap", "example_object_transport"]
STEPS = (1, 10, 50, 100)


def run(name, sort_every_step):
    p, o = prs.load_cfg(os.path.join(ROOT, "examples", name + ".cfg"))
    sim = prs.Simulation(p, 64.0, prs.BACKEND_EXTERNAL, ob.REFCUDA_PATH)
    sim.srand(p.seed)
    sim.reset()
    out = {"pos0": sim.get(prs.POSITION), "rad0": sim.get(prs.RADII)}
    si = o.timestep if sort_every_step else o.sort_interval
    for k in range(1, max(STEPS) + 1):
        sim.update(o.timestep, si)
        if k in STEPS:
            cs, ce = sim.get(prs.CELLSTART), sim.get(prs.CELLEND)
            occ = np.nonzero(cs != 0xFFFFFFFF)[0].astype(np.uint32)
            for key, arr in (("pos", sim.get(prs.POSITION)), ("vel", sim.get(prs.VELOCITY)), ("rad", sim.get(prs.RADII)),
                             ("phase", sim.get(prs.PHASE)), ("dead", sim.get(prs.DEAD)), ("hash", sim.get(prs.HASH)),
                             ("index", sim.get(prs.INDEX)), ("fr", sim.get(prs.ABSFORCE_R)), ("fa", sim.get(prs.ABSFORCE_A)),
                             ("occ", occ), ("cs_occ", cs[occ]), ("ce_occ", ce[occ])):
                out[f"{key}_{k}"] = arr
    sim.close()
    return out


if __name__ == "__main__":
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    prs.lib().cudaInit(0, None)
    for name in CFGS:
        for every in (False, True):
            tag = "sortall" if every else "refcadence"
            np.savez_compressed(os.path.join(dst, f"{name}.{tag}.npz"), **run(name, every))
            print("wrote", name, tag)
