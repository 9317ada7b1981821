"""Multi-GPU parity (-m gpu, needs >= 2 devices): the slab-decomposed run over NCCL must be
bit-equal to the single-GPU fused path on the same swarm (same kernels, same summation order)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

import particlerobotsimulations_b200 as prs
from particlerobotsimulations_b200 import multigpu
from tests import util

pytestmark = pytest.mark.gpu

NX, NY, PITCH, STEPS = multigpu.SELFCHECK["nx"], multigpu.SELFCHECK["ny"], multigpu.SELFCHECK["pitch"], 60


def _config():
    """the self-check swarm of multigpu.selfcheck_vs_single_gpu (bench.py runs the same check on its live ranks):
    150 dead robots drawn on step 0 (every rank draws the same global ids), velocities that make robots migrate"""
    return multigpu.selfcheck_config(os.path.join(util.EXAMPLES, "example.cfg"))


_initial_velocity = multigpu.selfcheck_velocity


def _worker(rank, world, port, out_path, bin_mode, exchange):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    p, o, geom = _config()
    prs.lib().prs_bin_set_mode(bin_mode)      # 0 auto, 1 onesweep + in-cell insertion sort, 2 cell binning
    sim = multigpu.make_hex_slab(p, o, geom, multigpu.CudaBackend, rank, world, dev, 5555, 0.01 * p.max_radius, exchange=exchange)
    n0 = sim.n
    sim.s.vel[:n0] = torch.from_numpy(_initial_velocity(sim.s.gid[:n0].cpu().numpy())).to(dev)
    snaps = {}
    for k in range(1, STEPS + 1):
        sim.step(o.timestep, o.timestep)
        if k in (1, 10, STEPS):
            snaps[k] = sim.gather_global(NX * NY)
    sim.check()      # capacity / "crossed two slabs" flags raised on the device
    stats = [None] * world
    dist.all_gather_object(stats, (sim.n, sim.stats["migrated"], sim.stats["halo"]))
    if rank == 0:
        np.savez(out_path, stats=np.array(stats), **{f"{key}_{k}": v for k, g in snaps.items() for key, v in g.items()})
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,bin_mode,exchange", [(2, 0, "p2p"), (2, 1, "nccl"), (2, 2, "p2p"), (2, 2, "nccl"), (4, 0, "p2p"),
                                                     (8, 0, "p2p"), (8, 2, "nccl")])
def test_slabs_bit_equal_to_single_gpu(world, bin_mode, exchange, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "slabs.npz")
    mp.spawn(_worker, args=(world, _free_port(), out, bin_mode, exchange), nprocs=world, join=True)
    got = np.load(out)
    p, o, geom = _config()
    torch.cuda.set_device(0)
    lib = prs.lib()
    lib.prs_set_stream(None)
    sim = prs.Simulation(p, 64.0, prs.BACKEND_FUSED)
    sim.srand(p.seed)                # main.cpp:929 — the dead draw continues this stream
    sim.init_hex(NX, NY, PITCH, 0.01 * p.max_radius, 5555)
    sim.set(prs.VELOCITY, _initial_velocity(np.arange(NX * NY)))
    stats = got["stats"]
    assert stats[:, 0].sum() == NX * NY and stats[:, 1].sum() > 0 and stats[:, 2].sum() > 0
    for k in range(1, STEPS + 1):
        sim.update(o.timestep, o.timestep)
        if k in (1, 10, STEPS):
            assert np.array_equal(got[f"pos_{k}"], sim.get(prs.POSITION)), k
            assert np.array_equal(got[f"vel_{k}"], sim.get(prs.VELOCITY)), k
            assert np.array_equal(got[f"rad_{k}"], sim.get(prs.RADII)), k
            assert np.array_equal(got[f"phase_{k}"], sim.get(prs.PHASE)), k
    sim.close()
