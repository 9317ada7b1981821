#!/usr/bin/env python
"""compute-sanitizer target: a 65 536-robot hex block through the fused path for a few steps with each collide variant
(thread per robot + dense start table, patch kernel), sort every step and at the reference cadence, plus one example cfg."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particlerobotsimulations_b200 as prs
lib = prs.lib()
lib.cudaInit(0, None)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for tile, dense, warp_max in ((0, 1, 0), (1, 0, 0), (0, 0, 16384)):
    lib.prs_set_collide_tile(tile); lib.prs_set_collide_dense(dense); lib.prs_set_collide_warp_max(warp_max)
    for cadence in (1, 0):
        p, o = prs.load_cfg(os.path.join(root, "examples", "example.cfg"))
        p.nCells = 256 * 256
        lib.prs_params_set_world(C.byref(p), 512, 64.0)
        sim = prs.Simulation(p, 64.0, prs.BACKEND_FUSED)
        sim.init_hex(256, 256, 0.17, 0.01 * p.max_radius, 5555)
        for k in range(8):
            sim.update(o.timestep, o.timestep if cadence else o.sort_interval)
            if k == 3:
                sim.sync()
        sim.sync()
        sim.close()
    p, o = prs.load_cfg(os.path.join(root, "examples", "example_gap.cfg"))
    sim = prs.Simulation(p, 64.0, prs.BACKEND_FUSED)
    sim.srand(p.seed); sim.reset()
    for k in range(6):
        sim.update(o.timestep, o.timestep)
    sim.sync(); sim.close()
print("sanitize_probe done")
