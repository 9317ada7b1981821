"""Diagnostic (not a test): stability of a synthetic hex swarm under the reference physics.

    python tests/diag_swarm.py <nx> <ny> <pitch> <steps> <every> <backend: fused|ref> [oracle]

Prints NaN count, fastest robot and fullest cell every <every> steps.  Used to show that the hex
lattice at pitch 2*min_radius = 0.155 (BASELINE.md S1 as first specified) is numerically unstable
in the reference's own DEM model — with the reference's own kernels too — which is why bench.py
uses pitch 0.17 (see DESIGN.md "workload").
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particlerobotsimulations_b200 as prs  # noqa: E402
from oracle import binding as ob  # noqa: E402
import bench  # noqa: E402

nx, ny, pitch, steps, every, backend = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
with_oracle = len(sys.argv) > 7
p, o, geom = bench.swarm_config(prs, 20, world64=(backend == "ref"), nx=nx, ny=ny, pitch=pitch)
print(geom, flush=True)
lib = prs.lib()
lib.cudaInit(0, None)
if backend == "ref":
    sim = prs.Simulation(p, geom["half"], prs.BACKEND_EXTERNAL, ob.REFCUDA_PATH)
else:
    sim = prs.Simulation(p, geom["half"], prs.BACKEND_FUSED)
sim.init_hex(nx, ny, pitch, bench.JITTER_FRAC * p.max_radius, bench.SEED)
s = None
if with_oracle:
    L = ob.lib()
    L.prso_set_threads(L.prso_get_max_threads())
    s = ob.OracleSim(p, geom["half"])
    s.view("pos")[:] = sim.get(prs.POSITION)
    s.view("rad")[:] = p.min_radius
for k in range(1, steps + 1):
    sim.update(o.timestep, o.timestep)
    if s:
        s.update(o.timestep, o.timestep)
    if k % every == 0 or k == 1:
        gp, gv, r = sim.get(prs.POSITION), sim.get(prs.VELOCITY), sim.get(prs.RADII)
        cs, ce = sim.get(prs.CELLSTART), sim.get(prs.CELLEND)
        m = cs != 0xFFFFFFFF
        msg = (f"{k}: nan {int(np.isnan(gp).sum())} vmax {np.nanmax(np.abs(gv)):.3g} moving {(np.abs(gv).max(1) > 0).mean():.3f} "
               f"oscillating {(r > p.min_radius * 1.001).mean():.3f} fullest cell {int((ce[m] - cs[m]).max())}")
        if s:
            msg += f" | oracle vmax {np.nanmax(np.abs(s.view('vel'))):.3g} hash eq {bool(np.array_equal(sim.get(prs.HASH), s.view('hash')))}"
        print(msg, flush=True)
