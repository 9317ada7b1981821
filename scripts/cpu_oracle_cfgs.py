#!/usr/bin/env python
"""CPU-1T row of BASELINE.md §2 for the example cfgs: the single-thread C++ oracle, 1000 steps, reference cadence."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particlerobotsimulations_b200 as prs
from oracle import binding as ob
ob.build()
L = ob.lib()
L.prso_set_threads(1)
for name in ["example", "example_dead_cells", "example_obstacle", "example_gap", "example_object_transport"]:
    p, o = prs.load_cfg(os.path.join(os.path.dirname(__file__), "..", "examples", name + ".cfg"))
    s = ob.OracleSim(p)
    s.srand(p.seed); s.reset()
    for _ in range(10): s.update(o.timestep, o.sort_interval)
    t0 = time.perf_counter()
    for _ in range(1000): s.update(o.timestep, o.sort_interval)
    dt = time.perf_counter() - t0
    print(f"CPU-1T oracle  {name:28s} N={p.nCells:5d}  {1e3*dt:8.1f} us/step  {1000/dt:9.1f} steps/s")
    s.close()
