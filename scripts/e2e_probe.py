"""S1 host-buffer step (prs_sim_update_host) with 1 / 2 / 4 / 8 chunks of the pipelined plan: ms per step."""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import particlerobotsimulations_b200 as prs  # noqa: E402

lib = prs.lib()
lib.cudaInit(0, None)
p, o, geom = bench.swarm_config(prs, 20)
sim = prs.Simulation(p, geom["half"], prs.BACKEND_FUSED)
sim.init_hex(geom["nx"], geom["ny"], geom["pitch"], 0.01 * p.max_radius, 5555)
n = int(p.nCells)
for _ in range(260):
    sim.update(o.timestep, o.timestep)
h = [torch.empty(s, dtype=torch.float32).pin_memory() for s in ((n, 2), (n, 2), (n,))]
for which, t in zip((prs.POSITION, prs.VELOCITY, prs.RADII), h):
    lib.prs_sim_get(sim._h, which, t.data_ptr(), t.numel() * 4)
out = {}
for chunks in (1, 2, 4, 8, 16, 4):
    lib.prs_set_plan_chunks(chunks)
    def step():
        lib.prs_sim_update_host(sim._h, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(),
                                o.timestep, o.timestep)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        step()
    torch.cuda.synchronize()
    out[f"chunks_{chunks}"] = round(1e3 * (time.perf_counter() - t0) / 100, 4)
    out[f"binned_{chunks}"] = lib.prs_bin_active()
print(json.dumps(out))
