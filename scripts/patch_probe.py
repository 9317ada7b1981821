#!/usr/bin/env python
"""Tuning aid: S1 at the timed state, collide stage time with the patch kernel on / off, patch statistics.
usage: patch_probe.py [--log2 20] [--rows 8] [--steps 20] [--evolve 260] [--pitch 0.17]"""
import argparse, ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import particlerobotsimulations_b200 as prs

ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, default=20)
ap.add_argument("--rows", type=int, default=8)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--evolve", type=int, default=260)
ap.add_argument("--pitch", type=float, default=None)
ap.add_argument("--only", type=int, default=None, help="only tile = this value")
ap.add_argument("--k1x2", type=int, default=1, help="K1 with two robots per thread (1) or one (0)")
ap.add_argument("--dense", type=int, default=1, help="collide reads the dense start table (1) or cellStart / cellEnd (0)")
a = ap.parse_args()
lib = prs.lib()
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
lib.prs_set_stream(C.c_void_p(stream.cuda_stream))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for tile in ([a.only] if a.only is not None else [0, 1]):
    lib.prs_set_collide_tile(tile)
    lib.prs_set_patch_rows(a.rows)
    lib.prs_set_k1_x2(a.k1x2)
    lib.prs_set_collide_dense(a.dense)
    p, o, geom = bench.swarm_config(prs, a.log2, pitch=a.pitch)
    sim = prs.Simulation(p, geom["half"], prs.BACKEND_FUSED)
    sim.init_hex(geom["nx"], geom["ny"], geom["pitch"], bench.JITTER_FRAC * p.max_radius, bench.SEED)
    for k in range(a.evolve):
        sim.update(o.timestep, o.timestep)
        if k == 3:
            sim.sync()
    torch.cuda.synchronize()
    ms = bench.timed_steps(torch, sim, o.timestep, o.timestep, a.steps, 3, flush, stat="mean")
    lib.prs_stage_timing(1)
    lib.prs_patch_stats(1, None)
    for _ in range(a.steps):
        flush.fill_(1)
        sim.update(o.timestep, o.timestep)
    t = (C.c_float * 6)(); c = (C.c_uint * 6)()
    lib.prs_stage_times(t, c)
    lib.prs_stage_timing(0)
    st = (C.c_uint * 3)()
    lib.prs_patch_stats(0, st)
    out[f"tile{tile}"] = dict(ms_per_step=ms, stages_us={bench.STAGES[i]: 1e3 * t[i] / c[i] for i in range(6) if c[i]},
                              patches_per_step=[x / a.steps for x in st], vel_crc=int(np.bitwise_xor.reduce(sim.get(prs.VELOCITY).view(np.uint32).ravel())))
    sim.close()
print(json.dumps(out))
