#!/usr/bin/env python
"""compute-sanitizer --tool racecheck target: the patch kernel (shared-memory lists, barriers, TMA landing zone) on a small hex block."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particlerobotsimulations_b200 as prs
lib = prs.lib()
lib.cudaInit(0, None)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib.prs_set_collide_tile(1); lib.prs_set_collide_warp_max(0)
p, o = prs.load_cfg(os.path.join(root, "examples", "example.cfg"))
p.nCells = 96 * 96
lib.prs_params_set_world(C.byref(p), 512, 64.0)
sim = prs.Simulation(p, 64.0, prs.BACKEND_FUSED)
sim.init_hex(96, 96, 0.17, 0.01 * p.max_radius, 5555)
for k in range(3):
    sim.update(o.timestep, o.timestep)
sim.sync(); sim.close()
print("racecheck_probe done")
