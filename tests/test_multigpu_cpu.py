"""Multi-rank host logic on CPU (world_size 2 and 3, gloo): the slab decomposition, migration and
halo exchange of particlerobotsimulations_b200/multigpu.py driven with a stand-in compute backend
built on the CPU oracle, compared with a single-process oracle run of the same swarm."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import particlerobotsimulations_b200 as prs
from particlerobotsimulations_b200 import multigpu
from oracle import binding as ob
from tests import util

NX = NY = 48
PITCH = 0.17
STEPS = 40


class OracleBackend:
    """CPU stand-in for multigpu.CudaBackend (same call surface) on CPU torch tensors."""

    def __init__(self, params, world_half):
        self.p, self.half, self.L = params, world_half, ob.lib()
        n = int(params.nCells)
        self.all_rng = (ob.RngState * n)()
        self.L.prso_curand_setup(self.all_rng, params.seed, n)

    @staticmethod
    def _a(t):
        return t.numpy()

    def k1(self, s, time, dt, n, do_hash):
        P, L = C.byref(self.p), self.L
        if time >= 0:
            L.prso_update_rad(P, s.fa.numpy().ctypes.data, s.fr.numpy().ctypes.data, s.rad.numpy().ctypes.data,
                              s.phase.numpy().ctypes.data, time, dt, s.dead.numpy().ctypes.data, n)
        L.prso_integrate(P, s.pos.numpy().ctypes.data, s.vel.numpy().ctypes.data, s.rad.numpy().ctypes.data, dt, n, self.half)
        if do_hash:
            L.prso_calc_hash(P, s.pos.numpy().ctypes.data, s.hash.numpy().ctypes.data, s.index.numpy().ctypes.data, n)

    def sort(self, keys_in, keys_out, vals_out, n, gid):
        k = keys_in.numpy()[:n]
        order = np.lexsort((gid.numpy()[:n], k))      # by hash, ties by global id
        keys_out.numpy()[:n] = k[order]
        vals_out.numpy()[:n] = order.astype(np.int32)

    def gather(self, pr, svel, index, s, n):
        idx = index.numpy()[:n]
        out = pr.numpy()
        out[:n, 0:2] = s.pos.numpy()[idx]
        out[:n, 2] = s.rad.numpy()[idx]
        out[:n, 3] = idx.astype(np.uint32).view(np.float32)
        svel.numpy()[:n] = s.vel.numpy()[idx]

    def cell_table(self, cs, ce, hash_cat, n, slot0, cell_lo, ncells):
        cs_, ce_, h = cs.numpy(), ce.numpy(), hash_cat.numpy()[:n]
        cs_[cell_lo:cell_lo + ncells] = -1
        if n == 0:
            return
        first = np.nonzero(np.r_[True, h[1:] != h[:-1]])[0]
        cs_[h[first]] = slot0 + first
        ce_[h[first[1:] - 1]] = slot0 + first[1:]
        ce_[h[-1]] = slot0 + n

    def lower_bounds(self, hash_sorted, n, bounds, out):
        out.numpy()[:] = np.searchsorted(hash_sorted.numpy()[:n], bounds.numpy(), "left").astype(np.int32)

    def collide(self, s, pr, svel, cs, ce, k_begin, k_end, dt):
        prn = pr.numpy()                                  # whole buffer: the upper halo lies beyond k_end
        spos = np.ascontiguousarray(prn[:, 0:2])
        srad = np.ascontiguousarray(prn[:, 2])
        cap = s.vel.shape[0]
        index = np.full(k_end, cap, np.uint32)          # lower-halo slots are computed into a dump entry
        index[k_begin:] = prn[k_begin:k_end, 3].copy().view(np.uint32)
        vel = np.zeros((cap + 1, 2), np.float32)
        fa, fr = np.zeros(cap + 1, np.float32), np.zeros(cap + 1, np.float32)
        fr[:cap] = s.fr.numpy()
        sv = np.ascontiguousarray(svel.numpy())
        self.L.prso_collide(C.byref(self.p), vel.ctypes.data, fa.ctypes.data, fr.ctypes.data, spos.ctypes.data, sv.ctypes.data,
                            srad.ctypes.data, index.ctypes.data, cs.numpy().ctypes.data, ce.numpy().ctypes.data, k_end, dt)
        own = index[k_begin:]
        s.vel.numpy()[own] = vel[own]
        s.fa.numpy()[own] = fa[own]
        s.fr.numpy()[own] = fr[own]

    def min_light_distance(self, pos, n, out):
        out.numpy()[0] = self.L.prso_min_light_distance(C.byref(self.p), pos.numpy().ctypes.data, n) if n else 3e38

    def update_phase(self, pos, phase, spacing, min_d, n):
        self.L.prso_update_phase(C.byref(self.p), pos.numpy().ctypes.data, phase.numpy().ctypes.data, spacing, float(min_d[0]), n)

    def rng_setup(self, rng, gid, n):
        words = np.frombuffer(bytes(self.all_rng), np.int32).reshape(-1, 12)
        rng.numpy()[:n] = words[gid.numpy()[:n]]

    def add_noise(self, rng, phase, std, n):
        self.L.prso_add_normal_noise(rng.numpy().ctypes.data, phase.numpy().ctypes.data, std, n)


def _config():
    p, o = util.cfg("example")
    p.nCells = NX * NY
    p.light_x, p.light_y = -6.0, 0.0
    geom = dict(nx=NX, ny=NY, pitch=PITCH, half=64.0)
    return p, o, geom


def _initial_velocity(gid):
    v = np.zeros((len(gid), 2), np.float32)
    v[:, 1] = (1.5 * np.sin(0.37 * gid.astype(np.float64))).astype(np.float32)
    return v


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    p, o, geom = _config()
    sim = multigpu.make_hex_slab(p, o, geom, OracleBackend, rank, world, torch.device("cpu"), 5555, 0.01 * p.max_radius)
    sim.s.vel[: sim.n] = torch.from_numpy(_initial_velocity(sim.s.gid[: sim.n].numpy()))   # makes robots cross slabs
    snaps = {}
    for k in range(1, STEPS + 1):
        sim.step(o.timestep, o.timestep)
        if k in (1, 5, STEPS):
            snaps[k] = sim.gather_global(NX * NY)
    if rank == 0:
        np.savez(out_path, migrated=sim.stats["migrated"], halo=sim.stats["halo"],
                 **{f"{key}_{k}": v for k, g in snaps.items() for key, v in g.items()})
    stats = [None] * world
    dist.all_gather_object(stats, (sim.n, sim.stats["migrated"], sim.stats["halo"]))
    if rank == 0:
        np.save(out_path + ".stats.npy", np.array(stats))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _single_process_reference():
    p, o, geom = _config()
    ids = np.arange(NX * NY)
    pos0 = multigpu.hex_block_positions(ids, NX, NY, PITCH, 0.01 * p.max_radius, 5555)
    s = ob.OracleSim(p, 64.0)
    s.view("pos")[:] = pos0
    s.view("rad")[:] = p.min_radius
    s.view("vel")[:] = _initial_velocity(ids)
    snaps = {}
    for k in range(1, STEPS + 1):
        s.update(o.timestep, o.timestep)
        if k in (1, 5, STEPS):
            snaps[k] = dict(pos=s.get("pos"), vel=s.get("vel"), rad=s.get("rad"), phase=s.get("phase"))
    return snaps, p


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_match_single_process(world, tmp_path):
    out = str(tmp_path / "slabs.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    stats = np.load(out + ".stats.npy")
    ref, p = _single_process_reference()
    assert stats[:, 0].sum() == NX * NY                   # every robot owned exactly once
    assert stats[:, 2].sum() > 0                          # halos were exchanged
    assert stats[:, 1].sum() > 0                          # robots migrated between slabs
    for k in (1, 5, STEPS):
        assert np.all(got[f"owner_{k}"] >= 0)
        # the phase noise stream belongs to the robot (seeded by global id): bit-equal at any time
        assert np.array_equal(got[f"phase_{k}"], ref[k]["phase"])
        vs = max(float(np.abs(ref[k]["vel"]).max()), 1e-3)
        # same arithmetic in the same order on every rank: bit-equal to the single-process run
        assert np.array_equal(got[f"pos_{k}"], ref[k]["pos"]), k
        assert np.array_equal(got[f"vel_{k}"], ref[k]["vel"]), k
        assert np.array_equal(got[f"rad_{k}"], ref[k]["rad"]), k
    # owners follow the rows: robots end up on the rank whose row range contains them
    rows = multigpu.slab_rows(p, NY, PITCH, world)
    r = multigpu.grid_row_of(got[f"pos_{STEPS}"][:, 1], p)
    expect = np.searchsorted(np.array(rows[1:]), r, "right")
    assert np.array_equal(got[f"owner_{STEPS}"], expect)


def test_hex_generator_matches_library_and_rows_partition():
    p, o, geom = _config()
    ids = np.arange(NX * NY)
    pos = multigpu.hex_block_positions(ids, NX, NY, PITCH, 0.01 * p.max_radius, 5555)
    assert pos.shape == (NX * NY, 2) and np.all(np.isfinite(pos))
    d = np.linalg.norm(pos[1] - pos[0])
    assert abs(d - PITCH) < 0.01
    for world in (2, 3, 8):
        rows = multigpu.slab_rows(p, NY, PITCH, world)
        assert rows[0] == 0 and rows[-1] == p.gridSize.y and all(b > a for a, b in zip(rows, rows[1:]))
        r = multigpu.grid_row_of(pos[:, 1], p)
        owner = np.searchsorted(np.array(rows[1:]), r, "right")
        counts = np.bincount(owner, minlength=world)
        assert counts.min() > 0.5 * NX * NY / world
