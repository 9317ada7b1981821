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


def _config(scenario="plain"):
    """the self-check swarm of multigpu.selfcheck_vs_single_gpu (bench.py runs the same check on its live ranks):
    150 dead robots drawn on step 0 (every rank draws the same global ids), velocities that make robots migrate.
    scenario "object": example_object_transport.cfg physics (nDead = -1: the last robot is the transported object,
    twice the radius, its own mass / friction / attraction factors) on the same lattice"""
    if scenario == "object":
        p, o, geom = multigpu.selfcheck_config(os.path.join(util.EXAMPLES, "example_object_transport.cfg"))
        p.nDead = -1
        return p, o, geom
    return multigpu.selfcheck_config(os.path.join(util.EXAMPLES, "example.cfg"))


WRAP_SHIFT = 42.3     # "wrap" scenario: the block sits under the top wall of the reference's world, where the row index wraps


def _object_start(n_total):
    """the object (robot n_total - 1) is dropped into the middle of the block, just below the cut between the two middle
    slabs, moving up: it ploughs through the lattice and changes owner"""
    return np.float32([0.031, -0.085]), np.float32([0.0, 3.0])


_initial_velocity = multigpu.selfcheck_velocity


def _worker(rank, world, port, out_path, bin_mode, exchange, scenario="plain"):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    p, o, geom = _config(scenario)
    prs.lib().prs_bin_set_mode(bin_mode)      # 0 auto, 1 onesweep + in-cell insertion sort, 2 cell binning
    if scenario == "object":
        # the object starts near the middle cut: the rank that owns that row must hold it from the start
        obj_pos, obj_vel = _object_start(NX * NY)
        rows = multigpu.slab_rows(p, NY, PITCH, world)
        ids = np.arange(NX * NY, dtype=np.int64)
        pos = multigpu.hex_block_positions(ids, NX, NY, PITCH, 0.01 * p.max_radius, 5555)
        pos[-1] = obj_pos
        r = multigpu.grid_row_of(pos[:, 1], p)
        keep = (r >= rows[rank]) & (r < rows[rank + 1])
        sim = multigpu.SlabSim(p, o, multigpu.CudaBackend(p, geom["half"]), rank, world, dev, pos[keep], ids[keep], rows,
                               capacity=int(NX * NY / world * 1.5) + 4096, halo_cap=16384, mig_cap=4096, exchange=exchange)
    elif scenario == "wrap":
        # The reference's own world: +-64 on 512 rows of 0.235 — rows 512.. of the robots above y = 56.3 alias rows 0.. (SURVEY.md
        # Q9).  The block is pushed up under the top wall; the slabs form a ring, the first one owns the aliased rows: robots that
        # cross y = 56.3 migrate from the LAST slab to the FIRST, and the two exchange halos across the seam.
        ids = np.arange(NX * NY, dtype=np.int64)
        pos = multigpu.hex_block_positions(ids, NX, NY, PITCH, 0.01 * p.max_radius, 5555)
        pos[:, 1] += np.float32(WRAP_SHIFT)
        r_un = multigpu.grid_row_of(pos[:, 1], p)                      # unwrapped: the top lattice rows lie beyond row 511
        r = r_un & (int(p.gridSize.y) - 1)
        lo, hi = int(r_un.min()), int(p.gridSize.y)
        rows = [0] + [lo + ((hi - lo) * (b + 1)) // world for b in range(world - 1)] + [hi]   # the first slab also owns rows 0.. (the seam)
        keep = (r >= rows[rank]) & (r < rows[rank + 1])
        sim = multigpu.SlabSim(p, o, multigpu.CudaBackend(p, geom["half"]), rank, world, dev, pos[keep], ids[keep], rows,
                               capacity=NX * NY + 4096, halo_cap=32768, mig_cap=8192, exchange=exchange, wrap=True)
    elif scenario == "rebalance":
        # lopsided cuts to start with: rebalance() (called twice below) has to move the boundaries and a large part of the swarm
        ids = np.arange(NX * NY, dtype=np.int64)
        pos = multigpu.hex_block_positions(ids, NX, NY, PITCH, 0.01 * p.max_radius, 5555)
        r = multigpu.grid_row_of(pos[:, 1], p)
        lo, hi = int(r.min()), int(r.max()) + 1
        rows = [0] + [lo + max(1, ((hi - lo) * (b + 1)) // (4 * world)) + b for b in range(world - 1)] + [int(p.gridSize.y)]
        keep = (r >= rows[rank]) & (r < rows[rank + 1])
        sim = multigpu.SlabSim(p, o, multigpu.CudaBackend(p, geom["half"]), rank, world, dev, pos[keep], ids[keep], rows,
                               capacity=NX * NY + 4096, halo_cap=32768, mig_cap=8192, exchange=exchange)
    else:
        sim = multigpu.make_hex_slab(p, o, geom, multigpu.CudaBackend, rank, world, dev, 5555, 0.01 * p.max_radius, exchange=exchange)
    n0 = sim.n
    vel0 = _initial_velocity(sim.s.gid[:n0].cpu().numpy())
    if scenario == "object":
        vel0[sim.s.gid[:n0].cpu().numpy() == NX * NY - 1] = _object_start(NX * NY)[1]
    sim.s.vel[:n0] = torch.from_numpy(vel0).to(dev)
    snaps = {}
    moved = 0
    for k in range(1, STEPS + 1):
        if scenario == "rebalance" and k in (12, 40):
            moved += sim.rebalance()["moved_out"]
        sim.step(o.timestep, o.timestep)
        if k in (1, 10, STEPS):
            snaps[k] = sim.gather_global(NX * NY)
    sim.check()      # capacity / "crossed two slabs" flags raised on the device
    stats = [None] * world
    dist.all_gather_object(stats, (sim.n, sim.stats["migrated"] + moved, sim.stats["halo"]))
    if rank == 0:
        np.savez(out_path, stats=np.array(stats), **{f"{key}_{k}": v for k, g in snaps.items() for key, v in g.items()})
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,bin_mode,exchange,scenario", [
    (2, 0, "p2p", "plain"), (2, 1, "nccl", "plain"), (2, 2, "p2p", "plain"), (2, 2, "nccl", "plain"), (2, 0, "p2p", "object"),
    (2, 1, "nccl", "object"), (2, 0, "p2p", "wrap"), (2, 1, "nccl", "wrap"), (4, 0, "p2p", "wrap"), (2, 0, "p2p", "rebalance"), (4, 0, "p2p", "rebalance"), (4, 0, "p2p", "plain"), (4, 0, "p2p", "object"), (8, 0, "p2p", "plain"), (8, 2, "nccl", "plain")])
def test_slabs_bit_equal_to_single_gpu(world, bin_mode, exchange, scenario, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    out = str(tmp_path / "slabs.npz")
    mp.spawn(_worker, args=(world, _free_port(), out, bin_mode, exchange, scenario), nprocs=world, join=True)
    got = np.load(out)
    p, o, geom = _config(scenario)
    torch.cuda.set_device(0)
    lib = prs.lib()
    lib.prs_set_stream(None)
    sim = prs.Simulation(p, 64.0, prs.BACKEND_FUSED)
    sim.srand(p.seed)                # main.cpp:929 — the dead draw continues this stream
    sim.init_hex(NX, NY, PITCH, 0.01 * p.max_radius, 5555)
    vel0 = _initial_velocity(np.arange(NX * NY))
    if scenario == "wrap":
        pos0 = sim.get(prs.POSITION)
        pos0[:, 1] += np.float32(WRAP_SHIFT)
        sim.set(prs.POSITION, pos0)
    if scenario == "object":
        obj_pos, obj_vel = _object_start(NX * NY)
        pos0 = sim.get(prs.POSITION)
        pos0[-1] = obj_pos
        sim.set(prs.POSITION, pos0)
        vel0[-1] = obj_vel
    sim.set(prs.VELOCITY, vel0)
    stats = got["stats"]
    assert stats[:, 0].sum() == NX * NY and stats[:, 1].sum() > 0 and stats[:, 2].sum() > 0
    for k in range(1, STEPS + 1):
        sim.update(o.timestep, o.timestep)
        if k in (1, 10, STEPS):
            assert np.array_equal(got[f"pos_{k}"], sim.get(prs.POSITION)), k
            assert np.array_equal(got[f"vel_{k}"], sim.get(prs.VELOCITY)), k
            assert np.array_equal(got[f"rad_{k}"], sim.get(prs.RADII)), k
            assert np.array_equal(got[f"phase_{k}"], sim.get(prs.PHASE)), k
            if scenario == "object":       # the object really is the heavy, never-oscillating robot on both sides
                assert got[f"rad_{k}"][-1] == np.float32(p.min_radius) * np.float32(p.radFactor)
    if scenario == "wrap":
        rows_end = multigpu.grid_row_of(got[f"pos_{STEPS}"][:, 1], p)
        assert (rows_end >= int(p.gridSize.y)).sum() > 100          # robots above the seam ...
        assert np.all(got[f"owner_{STEPS}"][rows_end >= int(p.gridSize.y)] == 0)   # ... live on the FIRST slab
    if scenario == "rebalance":
        assert stats[:, 1].sum() > NX * NY // 8 and stats[:, 0].max() < 1.25 * NX * NY / world, stats     # moved a lot, ended balanced
    if scenario == "object" and world == 2:
        rows = multigpu.slab_rows(p, NY, PITCH, world)
        first_owner = int(np.searchsorted(np.array(rows[1:]), multigpu.grid_row_of(_object_start(NX * NY)[0][1:2], p)[0], "right"))
        owners = [int(got[f"owner_{k}"][-1]) for k in (1, 10, STEPS)]
        assert any(w != first_owner for w in owners), (first_owner, owners)    # the object changed hands across the cut
    sim.close()


# ---- the C++ launcher: ParticleBot --gpus N (csrc/prs_multi.cpp), one process per rank, no Python on the path -----------------
def _final_state(path):
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw, np.uint64, 1)[0])
    f = np.frombuffer(raw, np.uint32, 6 * n, 8)
    return {"pos": f[:2 * n], "vel": f[2 * n:4 * n], "rad": f[4 * n:5 * n], "phase": f[5 * n:6 * n]}


def _run_launcher(tmp_path, cfg_text, ranks, steps, tag, extra=()):
    import subprocess
    cfg = tmp_path / f"{tag}.cfg"
    cfg.write_text(cfg_text)
    out = tmp_path / f"{tag}_{ranks}.bin"
    exe = os.path.join(util.ROOT, "particlerobotsimulations_b200", "ParticleBot")
    cmd = [exe, str(cfg), "--steps", str(steps), "--quiet", "--final-state", str(out)] + list(extra)
    if ranks > 1:
        cmd += ["--gpus", str(ranks)]
        if torch.cuda.device_count() < ranks:
            cmd.append("--oversubscribe")      # ranks share the device (time-sliced): slow, but the same code path
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return _final_state(out), r.stderr


HEX_CFG = """nCells
{n}
nDead
{dead}
light_x
-30
light_y
0
max_time
1e30
seed
5555
sort_interval
{sort}
csv_filename
launcher_{tag}.csv
testing
{testing}
dump_interval
0.1
init_config
hexblock
hexblock_nx
{nx}
hexblock_ny
{ny}
hexblock_pitch
0.17
hexblock_jitter
0.01
"""


@pytest.mark.gpu
@pytest.mark.parametrize("ranks,extra", [(2, ()), (3, ()), (2, ("--no-fused-exchange",)), (2, ("--no-overlap",))])
def test_cpp_launcher_hexblock_equals_one_gpu(tmp_path, ranks, extra):
    """`ParticleBot cfg --gpus N` (forked ranks, CUDA IPC mailboxes, shared-memory control plane) == `ParticleBot cfg` on one
    GPU, bit for bit, after 60 steps of a 49 152-robot block with 150 dead robots drawn at step 0 — positions, velocities,
    radii, phases in robot order — and the CSV of rank 0 equals the single-GPU CSV byte for byte (centroid summed in robot
    order).  On a box with fewer GPUs than ranks the ranks share the device."""
    text = HEX_CFG.format(n=256 * 192, dead=150, sort=0.01, tag="hex", testing=0, nx=256, ny=192)
    one, _ = _run_launcher(tmp_path, text, 1, 60, "hex")
    csv_one = (tmp_path / "launcher_hex.csv").read_bytes()
    many, err = _run_launcher(tmp_path, text, ranks, 60, "hex", extra)   # extra: one kernel per exchange operation / no overlap
    csv_many = (tmp_path / "launcher_hex.csv").read_bytes()
    assert f"on {ranks} ranks" in err
    for k in ("pos", "vel", "rad", "phase"):
        assert np.array_equal(one[k], many[k]), k
    assert csv_one == csv_many and csv_one.count(b"\n") >= 8


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["example.cfg", "example_dead_cells.cfg", "example_object_transport.cfg"])
def test_cpp_launcher_reference_cfgs_equal_one_gpu(tmp_path, name):
    """the reference's own cfgs (aggregation placement on the glibc stream, dead draw continuing it, the transported object,
    the reference's +-64 world whose hash wraps: ring of slabs) on 2 ranks, reference cadence, 300 steps, per-robot CSV"""
    text = open(os.path.join(util.ROOT, "examples", name)).read() + "\ntesting\n1\ndump_interval\n0.5\ncsv_filename\nlauncher_ref.csv\n"
    one, _ = _run_launcher(tmp_path, text, 1, 300, "ref")
    csv_one = (tmp_path / "launcher_ref.csv").read_bytes()
    two, _ = _run_launcher(tmp_path, text, 2, 300, "ref")
    csv_two = (tmp_path / "launcher_ref.csv").read_bytes()
    for k in ("pos", "vel", "rad", "phase"):
        assert np.array_equal(one[k], two[k]), k
    assert csv_one == csv_two and len(csv_one) > 1000
