"""Slab-decomposed multi-GPU execution of the particle-robot update.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  Because the cell
hash is row-major (hash = row * gridSize.x + column) a band of grid rows is a contiguous range of
the sorted arrays, so the world is cut into horizontal slabs of grid rows, one per rank:

  per step, on every rank
    1. [phase gate]  local min light distance -> all_reduce(MIN) -> phase offsets (+ XORWOW noise,
       the generator state travels with its robot and is seeded by GLOBAL id)
    2. controller + integrate (+ hash on sort steps) on the owned robots           (prs_slab_k1)
    3. [sort steps]  MIGRATION: robots whose new row left the slab are packed (full state record,
       88 B) and sent to the neighbour that owns the row; arrivals are appended     (isend/irecv)
    4. [sort steps]  local onesweep sort of (hash, local slot), ties put in global-id order
                                                                  (prs_slab_sort, prs_slab_fix_ties)
    5. gather into the packed sorted layout at a fixed offset                       (prs_slab_gather)
    6. HALO: the first/last HALO_ROWS grid rows of the sorted range are contiguous slices; they
       are sent to the lower/upper neighbour and received into the flanks            (isend/irecv)
    7. cell table over [lower halo | owned | upper halo]                            (prs_slab_cell_table)
    8. collide over the owned range, results scattered to local slots                (prs_slab_collide)

The data path has no collective: only neighbour sends of <= a few MB (latency-bound on NVLink),
one 4-byte all_reduce per phase update, and two small all_gathers of counts per step (the host
needs them to size the transfers — the two host syncs of the step).  Ownership changes only on
sort steps (the table is frozen between sorts), HALO_ROWS = 3 = the 2-row stencil +
1 guard row for drift between sorts.  Limits of this version: the hash wrap-around (Q9) is not
exchanged — the world must fit the grid — a robot may cross at most one slab per sort, object
transport (nDead == -1) is single-GPU only.  Robots of one cell are ordered by GLOBAL id after the
local sort (prs_slab_fix_ties), which is the order the reference's stable sort gives them, so the
forces are summed in the single-GPU order and the results are bit-equal for any number of slabs.

The compute calls go through a small backend object so that the CPU tests can drive the same
host logic with a stand-in (tests/test_multigpu_cpu.py injects one built on the oracle).
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist

HALO_ROWS = 3
RECORD_FIELDS = ("pos", "vel", "rad", "phase", "fa", "fr", "dead", "gid", "rng", "hash")


# --------------------------------------------------------------------------------------------------
def hex_block_positions(ids, nx, ny, pitch, jitter, seed):
    """Positions of robots `ids` (global ids, row-major lattice index) of the synthetic hex block —
    the same generator as Particlebot::initHexBlock (counter hash of (seed, id) for the jitter)."""
    i = np.asarray(ids, dtype=np.uint64)
    ix, iy = (i % np.uint64(nx)).astype(np.float32), (i // np.uint64(nx))
    pitch32 = np.float32(pitch)
    row = pitch32 * np.float32(0.8660254037844386)
    x0 = np.float32(-0.5) * (np.float32(nx - 1) * pitch32 + np.float32(0.5) * pitch32)
    y0 = np.float32(-0.5) * np.float32(ny - 1) * row

    def mix(z):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))

    with np.errstate(over="ignore"):
        h = mix((np.uint64(seed) << np.uint64(32)) ^ i)
    jit = np.float32(jitter)
    jx = ((h & np.uint64(0xFFFFFF)).astype(np.float32) / np.float32(8388608.0) - np.float32(1.0)) * jit
    jy = (((h >> np.uint64(24)) & np.uint64(0xFFFFFF)).astype(np.float32) / np.float32(8388608.0) - np.float32(1.0)) * jit
    odd = np.where((iy & np.uint64(1)) == np.uint64(1), np.float32(0.5) * pitch32, np.float32(0.0)).astype(np.float32)
    x = x0 + ix * pitch32 + odd + jx
    y = y0 + iy.astype(np.float32) * row + jy
    return np.stack([x, y], 1).astype(np.float32)


def grid_row_of(y, params):
    """grid row of a y coordinate: floor((y - origin.y) / cell.y) in fp32, like the device hash"""
    oy, cy = np.float32(params.worldOrigin.y), np.float32(params.cellSize.y)
    return np.floor((np.asarray(y, np.float32) - oy) / cy).astype(np.int64)


def slab_rows(params, ny, pitch, world):
    """Row boundaries R[0..world]: rank r owns grid rows [R[r], R[r+1]).  Lattice rows are split
    evenly and each cut is mapped to the grid row it falls in."""
    row = np.float32(pitch) * np.float32(0.8660254037844386)
    y0 = np.float32(-0.5) * np.float32(ny - 1) * row
    R = [0]
    for b in range(1, world):
        yb = y0 + np.float32((ny * b) // world) * row
        R.append(int(grid_row_of(yb, params)))
    R.append(int(params.gridSize.y))
    return R


# --------------------------------------------------------------------------------------------------
class CudaBackend:
    """The slab building blocks of libparticlebot_b200.so on CUDA tensors (current torch stream)."""

    def __init__(self, params, world_half):
        import particlerobotsimulations_b200 as prs
        self.lib = prs.lib()
        self.lib.prs_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.lib.prs_set_world_half_extent(world_half)
        self.lib.setParameters(C.byref(params))

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr())

    def k1(self, s, time, dt, n, do_hash):
        self.lib.prs_slab_k1(self._p(s.pos), self._p(s.vel), self._p(s.rad), self._p(s.phase), self._p(s.fa),
                             self._p(s.fr), self._p(s.dead), self._p(s.hash), self._p(s.index), time, dt, n, int(do_hash))

    def sort(self, keys_in, keys_out, vals_out, n, gid):
        """(hash, local slot) sorted by hash, robots of one cell in ascending GLOBAL id"""
        self.lib.prs_slab_sort(self._p(keys_in), None, self._p(keys_out), self._p(vals_out), n, 1)
        self.lib.prs_slab_fix_ties(self._p(keys_out), self._p(vals_out), self._p(gid), n)

    def gather(self, pr, svel, index, s, n):
        self.lib.prs_slab_gather(self._p(pr), self._p(svel), self._p(index), self._p(s.pos), self._p(s.vel), self._p(s.rad), n)

    def cell_table(self, cs, ce, hash_cat, n, slot0, cell_lo, ncells):
        self.lib.prs_slab_cell_table(self._p(cs), self._p(ce), self._p(hash_cat), n, slot0, cell_lo, ncells)

    def lower_bounds(self, hash_sorted, n, bounds, out):
        self.lib.prs_slab_lower_bounds(self._p(hash_sorted), n, self._p(bounds), bounds.numel(), self._p(out))

    def collide(self, s, pr, svel, cs, ce, k_begin, k_end, dt):
        self.lib.prs_slab_collide(self._p(s.vel), self._p(s.fa), self._p(s.fr), self._p(pr), self._p(svel), self._p(cs),
                                  self._p(ce), k_begin, k_end, dt)

    def min_light_distance(self, pos, n, out):
        self.lib.prs_min_light_distance(self._p(pos), n, self._p(out))

    def update_phase(self, pos, phase, spacing, min_d, n):
        self.lib.prs_update_phase_dev(self._p(pos), self._p(phase), spacing, self._p(min_d), n)

    def rng_setup(self, rng, gid, n):
        self.lib.prs_curand_setup_ids(self._p(rng), self._p(gid), n)

    def add_noise(self, rng, phase, std, n):
        self.lib.add_normal_noise(self._p(rng), self._p(phase), std, n)


class _State:
    pass


class SlabSim:
    """One rank's slab of the swarm.  `backend` supplies the compute calls (CudaBackend or a test
    stand-in); `group` is the torch.distributed process group (None = default)."""

    def __init__(self, params, opt, backend, rank, world, device, pos, gid, rows, capacity=None, group=None):
        self.p, self.opt, self.be = params, opt, backend
        self.rank, self.world, self.dev, self.group = rank, world, device, group
        self.R_lo, self.R_hi = rows[rank], rows[rank + 1]
        self.GX, self.GY = int(params.gridSize.x), int(params.gridSize.y)
        self.n = int(len(gid))
        cap = capacity or int(self.n * 1.25) + 65536
        self.cap = cap
        self.halo_cap = max(65536, int(cap * 0.1))
        f32, i32 = torch.float32, torch.int32
        s = self.s = _State()
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
        s.pos, s.vel = z((cap, 2), f32), z((cap, 2), f32)
        s.rad, s.phase, s.fa, s.fr = z(cap, f32), z(cap, f32), z(cap, f32), z(cap, f32)
        s.dead, s.gid, s.hash, s.index = z(cap, i32), z(cap, i32), z(cap, i32), z(cap, i32)
        s.rng = z((cap, 12), i32)
        s.pos[: self.n] = torch.from_numpy(np.ascontiguousarray(pos)).to(device)
        s.gid[: self.n] = torch.from_numpy(np.ascontiguousarray(gid).astype(np.int32)).to(device)
        s.rad[: self.n] = float(np.float32(params.min_radius))
        self.hash_sorted, self.index_sorted = z(cap, i32), z(cap, i32)
        ncat = cap + 2 * self.halo_cap
        self.pr, self.svel, self.hash_cat = z((ncat, 4), f32), z((ncat, 2), f32), z(ncat, i32)
        self.cs = torch.full((int(params.numCells),), -1, dtype=i32, device=device)
        self.ce = z(int(params.numCells), i32)
        self.min_d = z(16, f32)
        self.bounds = torch.tensor([min(self.R_lo + HALO_ROWS, self.GY) * self.GX, max(self.R_hi - HALO_ROWS, 0) * self.GX],
                                   dtype=i32, device=device)
        self.bounds_out = z(2, i32)
        self.time = np.float32(0.0)
        self.n_lo = self.n_hi = 0
        self.sorted_once = False
        self.stats = dict(migrated=0, halo=0)
        self.be.rng_setup(s.rng, s.gid, self.n)

    # ---- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _gate(time, interval, dt):
        t, T, d = np.float32(time), np.float32(interval), np.float32(dt)
        return bool(t - T * np.floor(t / T) < d)

    def _neigh(self):
        return (self.rank - 1 if self.rank > 0 else None), (self.rank + 1 if self.rank < self.world - 1 else None)

    def _all_counts(self, a, b):
        """every rank's (a, b): one small all_gather and the host sync that sizes the transfers"""
        mine = torch.tensor([a, b], dtype=torch.int64, device=self.dev)
        out = [torch.zeros(2, dtype=torch.int64, device=self.dev) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return [tuple(int(v) for v in t.tolist()) for t in out]

    def _exchange(self, send_dn, send_up, recv_dn, recv_up):
        """lists of tensors to/from the lower (dn) and upper (up) neighbour, posted as one batch"""
        dn, up = self._neigh()
        ops = []
        for t in send_dn:
            ops.append(dist.P2POp(dist.isend, t, dn, group=self.group))
        for t in send_up:
            ops.append(dist.P2POp(dist.isend, t, up, group=self.group))
        for t in recv_dn:
            ops.append(dist.P2POp(dist.irecv, t, dn, group=self.group))
        for t in recv_up:
            ops.append(dist.P2POp(dist.irecv, t, up, group=self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    # ---- migration (sort steps) --------------------------------------------------------------------
    def _migrate(self):
        s, n = self.s, self.n
        rows = torch.div(s.hash[:n], self.GX, rounding_mode="floor")
        down, up = rows < self.R_lo, rows >= self.R_hi      # empty on the edge ranks by construction
        dn_rank, up_rank = self._neigh()
        idx_dn, idx_up = torch.nonzero(down).flatten(), torch.nonzero(up).flatten()
        counts = self._all_counts(idx_dn.numel(), idx_up.numel())
        n_from_dn = counts[dn_rank][1] if dn_rank is not None else 0
        n_from_up = counts[up_rank][0] if up_rank is not None else 0
        if not any(c[0] or c[1] for c in counts):
            return
        fields = [s.pos, s.vel, s.rad, s.phase, s.fa, s.fr, s.dead, s.gid, s.rng, s.hash]
        send_dn = [f[idx_dn].contiguous() for f in fields] if idx_dn.numel() else []
        send_up = [f[idx_up].contiguous() for f in fields] if idx_up.numel() else []
        recv_dn = [torch.empty((n_from_dn,) + tuple(f.shape[1:]), dtype=f.dtype, device=self.dev) for f in fields] if n_from_dn else []
        recv_up = [torch.empty((n_from_up,) + tuple(f.shape[1:]), dtype=f.dtype, device=self.dev) for f in fields] if n_from_up else []
        self._exchange(send_dn, send_up, recv_dn, recv_up)
        n_leave = idx_dn.numel() + idx_up.numel()
        n_new = n - n_leave + n_from_dn + n_from_up
        assert n_new <= self.cap, "slab capacity exceeded"
        if n_leave:
            keep = torch.nonzero(~(down | up)).flatten()
        for k, f in enumerate(fields):
            parts = [f[keep] if n_leave else f[:n]]
            if n_from_dn:
                parts.append(recv_dn[k])
            if n_from_up:
                parts.append(recv_up[k])
            if len(parts) > 1 or n_leave:
                f[:n_new] = torch.cat(parts) if len(parts) > 1 else parts[0]
        # a robot may cross one slab at most: every arrival must now sit in this slab's rows
        if n_from_dn or n_from_up:
            r2 = torch.div(s.hash[n - n_leave:n_new], self.GX, rounding_mode="floor")
            assert bool(((r2 >= self.R_lo) & (r2 < self.R_hi)).all()), "a robot crossed more than one slab in one step"
        self.stats["migrated"] += n_leave
        self.n = n_new

    # ---- halo ----------------------------------------------------------------------------------------
    def _halo(self):
        n, HC = self.n, self.halo_cap
        dn_rank, up_rank = self._neigh()
        # slots of the first / last HALO_ROWS rows of the owned sorted range
        self.be.lower_bounds(self.hash_sorted, n, self.bounds, self.bounds_out)
        b0, b1 = (int(v) for v in self.bounds_out.tolist())
        k_dn = b0 if dn_rank is not None else 0
        k_up = (n - b1) if up_rank is not None else 0
        counts = self._all_counts(k_dn, k_up)
        n_lo = counts[dn_rank][1] if dn_rank is not None else 0
        n_hi = counts[up_rank][0] if up_rank is not None else 0
        assert n_lo <= HC and n_hi <= HC, "halo capacity exceeded"
        own = slice(HC, HC + n)
        self.hash_cat[own] = self.hash_sorted[:n]
        send_dn = [self.pr[HC:HC + k_dn], self.svel[HC:HC + k_dn], self.hash_cat[HC:HC + k_dn]] if k_dn else []
        send_up = [self.pr[HC + n - k_up:HC + n], self.svel[HC + n - k_up:HC + n], self.hash_cat[HC + n - k_up:HC + n]] if k_up else []
        recv_dn = [self.pr[HC - n_lo:HC], self.svel[HC - n_lo:HC], self.hash_cat[HC - n_lo:HC]] if n_lo else []
        recv_up = [self.pr[HC + n:HC + n + n_hi], self.svel[HC + n:HC + n + n_hi], self.hash_cat[HC + n:HC + n + n_hi]] if n_hi else []
        self._exchange(send_dn, send_up, recv_dn, recv_up)
        self.n_lo, self.n_hi = n_lo, n_hi
        self.stats["halo"] += n_lo + n_hi

    # ---- one step (Particlebot::update, particlebot.cpp:170-300, cut at the exchanges) ----------------
    def step(self, dt, sort_interval):
        p, s, be = self.p, self.s, self.be
        time = self.time
        phase_step = self._gate(time, p.phase_update_interval, dt)
        sort_step = self._gate(time, sort_interval, dt) or not self.sorted_once
        if phase_step:
            be.min_light_distance(s.pos, self.n, self.min_d)
            dist.all_reduce(self.min_d[:1], op=dist.ReduceOp.MIN, group=self.group)
            be.update_phase(s.pos, s.phase, 2.0 * float(np.float32(p.min_radius)), self.min_d, self.n)
            if p.phase_std:
                be.add_noise(s.rng, s.phase, float(p.phase_std), self.n)
        be.k1(s, float(time), float(dt), self.n, sort_step)
        if sort_step:
            self._migrate()
            be.sort(s.hash, self.hash_sorted, self.index_sorted, self.n, s.gid)
            self.sorted_once = True
        HC = self.halo_cap
        be.gather(self.pr[HC:], self.svel[HC:], self.index_sorted, s, self.n)
        self._halo()
        start = HC - self.n_lo
        n_tot = self.n_lo + self.n + self.n_hi
        row_lo, row_hi = max(self.R_lo - HALO_ROWS, 0), min(self.R_hi + HALO_ROWS, self.GY)
        be.cell_table(self.cs, self.ce, self.hash_cat[start:], n_tot, start, row_lo * self.GX, (row_hi - row_lo) * self.GX)
        be.collide(s, self.pr, self.svel, self.cs, self.ce, HC, HC + self.n, float(dt))
        self.time = np.float32(time + np.float32(dt))

    # ---- assembling global arrays (tests, observables) -------------------------------------------------
    def gather_global(self, n_total):
        """(pos, vel, rad, phase) of the whole swarm in global-id order on every rank (small swarms / tests)"""
        s, n = self.s, self.n
        local = dict(gid=s.gid[:n].cpu().numpy(), pos=s.pos[:n].cpu().numpy(), vel=s.vel[:n].cpu().numpy(),
                     rad=s.rad[:n].cpu().numpy(), phase=s.phase[:n].cpu().numpy())
        parts = [None] * self.world
        dist.all_gather_object(parts, local, group=self.group)
        out = dict(pos=np.zeros((n_total, 2), np.float32), vel=np.zeros((n_total, 2), np.float32),
                   rad=np.zeros(n_total, np.float32), phase=np.zeros(n_total, np.float32), owner=np.full(n_total, -1))
        for r, part in enumerate(parts):
            g = part["gid"]
            for k in ("pos", "vel", "rad", "phase"):
                out[k][g] = part[k]
            out["owner"][g] = r
        return out


def make_hex_slab(params, opt, geom, backend_factory, rank, world, device, seed, jitter, group=None):
    """Builds rank `rank`'s slab of the nx*ny hex block: every rank generates only the lattice rows
    around its slab and keeps the robots whose grid row it owns."""
    nx, ny, pitch = geom["nx"], geom["ny"], geom["pitch"]
    rows = slab_rows(params, ny, pitch, world)
    iy_lo = max((ny * rank) // world - 4, 0)
    iy_hi = min((ny * (rank + 1)) // world + 4, ny)
    ids = np.arange(iy_lo * nx, iy_hi * nx, dtype=np.int64)
    pos = hex_block_positions(ids, nx, ny, pitch, jitter, seed)
    r = grid_row_of(pos[:, 1], params)
    keep = (r >= rows[rank]) & (r < rows[rank + 1])
    be = backend_factory(params, geom["half"])
    n_expected = (nx * ny) // world
    return SlabSim(params, opt, be, rank, world, device, pos[keep], ids[keep], rows,
                   capacity=int(n_expected * 1.25) + 65536, group=group)


# --------------------------------------------------------------------------------------------------
def bench_slabs(args, rank, world, local_rank):
    """bench.py --gpus N (N > 1): 2^26 robots (or --robots-log2) slab-decomposed over the ranks."""
    import json
    import os
    import sys
    import particlerobotsimulations_b200 as prs
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    log2n = args.robots_log2 or 26
    p, o, geom = bench.swarm_config(prs, log2n)
    n_total = int(p.nCells)
    sim = make_hex_slab(p, o, geom, CudaBackend, rank, world, dev, bench.SEED, bench.JITTER_FRAC * p.max_radius)
    lib = prs.lib()
    sort_interval = o.timestep if args.sort_interval is None else args.sort_interval
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        sim.step(o.timestep, sort_interval)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = bench.ClockSampler(local_rank)
    sampler.start()
    lib.prs_launch_count(1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for a, b in ev:
        flush.fill_(1)
        a.record(stream)
        sim.step(o.timestep, sort_interval)
        b.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    launches = int(lib.prs_launch_count(0))
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    clocks = sampler.stop()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    own = torch.tensor([sim.n, sim.stats["migrated"], sim.stats["halo"]], dtype=torch.int64, device=dev)
    owns = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(owns, own)
    finite = torch.tensor([int(torch.isfinite(sim.s.pos[: sim.n]).all())], device=dev)
    dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    if rank == 0:
        value = n_total * args.steps / (total_ms_max * 1e-3)
        peak, peak_src = bench.measured_peak()
        per_bytes, b_alg, passes = bench.algorithmic_bytes(p, sort_interval <= o.timestep)
        per_rank = [[int(v) for v in x.tolist()] for x in owns]
        steps_run = args.steps + args.warmup
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.robots_log2 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{geom['name']}: {n_total} robots, hex {geom['nx']}x{geom['ny']} pitch {geom['pitch']}, "
                                   f"world +-{geom['half']:g}, grid {geom['grid']}^2, {world} slabs of grid rows",
                       "sort_interval": "timestep (sort every step)", "collide_mode": "exact",
                       "l2": "flushed between timed steps (256 MiB write)", "halo_rows": HALO_ROWS},
            "e2e": None, "gpu_launches": launches, "clocks": clocks,
            "roofline_step": {"bound": "hbm", "alg_bytes_per_particle_step": b_alg, "radix_passes": passes,
                              "achieved": b_alg * value / 1e9, "peak": peak * world, "unit": "GB/s",
                              "frac": b_alg * value / 1e9 / (peak * world), "peak_source": peak_src + f" x {world} GPUs"},
            "slabs": {"robots_per_rank": [x[0] for x in per_rank],
                      "migrated_per_step_per_rank": [x[1] / steps_run for x in per_rank],
                      "halo_robots_per_step_per_rank": [x[2] / steps_run for x in per_rank]},
            "state_finite": bool(finite.item()),
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
