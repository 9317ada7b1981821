"""Slab-decomposed multi-GPU execution of the particle-robot update.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  Because the cell
hash is row-major (hash = row * gridSize.x + column) a band of grid rows is a contiguous range of
the sorted arrays, so the world is cut into horizontal slabs of grid rows, one per rank:

  per step, on every rank
    1. [phase gate]  local min light distance -> all_reduce(MIN) -> phase offsets (+ XORWOW noise,
       the generator state travels with its robot and is seeded by GLOBAL id)
    2. controller + integrate (+ hash on sort steps) on the owned robots           (prs_slab_k1)
    3. [sort steps]  MIGRATION: robots whose new row left the slab are packed (full state record,
       92 B) for the neighbour that owns the row, the survivors are compacted, arrivals appended
                                              (prs_slab_migrate_pack / exchange / _migrate_unpack)
    4. [sort steps]  local onesweep sort of (hash, local slot), ties put in global-id order
                                                                                   (prs_slab_sort)
    5. gather into the packed sorted layout at a fixed offset                    (prs_slab_gather)
    6. HALO: the first/last HALO_ROWS grid rows of the sorted range are contiguous slices; they
       are packed, sent to the lower/upper neighbour and unpacked into the flanks
                                                    (prs_slab_halo_pack / exchange / _halo_unpack)
    7. cell table over [lower halo | owned | upper halo]                      (prs_slab_cell_table)
    8. collide over the owned range, results scattered to local slots            (prs_slab_collide)

EVERY COUNT LIVES ON THE DEVICE (csrc/prs_slab.cuh): kernels are launched for the slab's capacity
and read the live robot / halo / migrant counts from a 16-word device array; the exchanges move
fixed-size buffers whose first word is the record count.  A step is therefore a pure stream of
kernel launches and neighbour sends/receives — no host synchronisation and no device-to-host
copy anywhere in it (overflow and "crossed two slabs" conditions set sticky error bits that
`check()` reads when the caller asks).  The data path has no collective: neighbour transfers of a
few MB (latency-bound on NVLink) and one 4-byte all_reduce per phase update.  Ownership changes
only on sort steps (the table is frozen between sorts); HALO_ROWS = 3 = the 2-row
stencil + 1 guard row for drift between sorts.  `wrap=True` makes the slabs a ring in the row index
(the cell hash wraps around the grid, Q9).  Limit: a robot may cross at most one slab per sort.
Object transport (nDead == -1) works across slabs: sorted and halo records carry the robot's GLOBAL
id as identity, so every rank recognises the object (robot nCells - 1) wherever it lives.  Robots of one cell are ordered by GLOBAL id
after the local sort, which is the order the reference's stable sort gives them, so the forces are
summed in the single-GPU order and the results are bit-equal for any number of slabs.

The compute calls go through a small backend object so that the CPU tests can drive the same
host logic with a stand-in (tests/test_multigpu_cpu.py injects one built on the oracle).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

import particlerobotsimulations_b200 as prs

HALO_ROWS = 3


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


def balanced_rows(params, y_local, world, group=None):
    """Row boundaries R[0..world] that give every rank about the same number of robots, for ANY swarm:
    each rank histograms the grid rows of the robots it currently holds (`y_local`), the histograms are
    summed over the ranks, and the cuts are placed on the cumulative count (slabs by equal
    robot count).  Needs an initialised process group when world > 1."""
    gy = int(params.gridSize.y)
    rows = np.clip(grid_row_of(y_local, params), 0, gy - 1)
    hist = torch.from_numpy(np.bincount(rows, minlength=gy).astype(np.int64))
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        hist = hist.to(dev)
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
        hist = hist.cpu()
    cum = np.cumsum(hist.numpy())
    total = int(cum[-1])
    R = [0]
    for b in range(1, world):
        cut = int(np.searchsorted(cum, total * b / world, "left")) + 1     # first row AFTER the b/world quantile
        R.append(min(max(cut, R[-1] + 1), gy - (world - b)))
    R.append(gy)
    return R


# --------------------------------------------------------------------------------------------------
class CudaBackend:
    """The slab engine of libparticlebot_b200.so on CUDA tensors (current torch stream)."""

    def __init__(self, params, world_half):
        self.lib = prs.lib()
        self.lib.prs_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.lib.prs_set_world_half_extent(world_half)
        self.lib.setParameters(C.byref(params))
        self.slab = None

    def bind(self, sim):
        """device pointers of `sim`'s tensors -> prs_slab"""
        s, sl = sim.s, prs.Slab()
        for name, t in (("pos", s.pos), ("vel", s.vel), ("rad", s.rad), ("phase", s.phase), ("absForce_a", s.fa),
                        ("absForce_r", s.fr), ("dead", s.dead), ("gid", s.gid), ("rng", s.rng), ("hash", s.hash),
                        ("scratch", s.scratch), ("sortedPR", sim.pr), ("sortedVel", sim.svel), ("hash_cat", sim.hash_cat),
                        ("index_sorted", sim.index_sorted), ("cellStart", sim.cs), ("cellEnd", sim.ce),
                        ("counts", sim.counts), ("lists", sim.lists)):
            setattr(sl, name, t.data_ptr())
        sl.cap, sl.halo_cap, sl.mig_cap = sim.cap, sim.halo_cap, sim.mig_cap
        sl.row_lo, sl.row_hi, sl.halo_rows = sim.R_lo, sim.R_hi, HALO_ROWS
        ring = bool(getattr(sim, "wrap", False)) and sim.world > 1
        sl.has_dn, sl.has_up = int(sim.rank > 0 or ring), int(sim.rank < sim.world - 1 or ring)
        sl.wrap = int(ring)
        self.slab = sl
        self._ref = C.byref(sl)

    @staticmethod
    def _p(t):
        """device address of a tensor, or a raw address (peer-mapped mailbox), or None"""
        if t is None:
            return None
        return C.c_void_p(t if isinstance(t, int) else t.data_ptr())

    def signal(self, remote_flag_dn, remote_flag_up, seq):
        self.lib.prs_slab_signal(self._p(remote_flag_dn), self._p(remote_flag_up), seq)

    def wait(self, local_flag_dn, local_flag_up, seq):
        self.lib.prs_slab_wait(self._ref, self._p(local_flag_dn), self._p(local_flag_up), seq)

    def rng_setup(self, n):
        self.lib.prs_slab_rng_setup(self._ref, n)

    def k1(self, time, dt, do_hash):
        self.lib.prs_slab_k1(self._ref, time, dt, int(do_hash))

    def migrate_pack(self, send_dn, send_up):
        self.lib.prs_slab_migrate_pack(self._ref, self._p(send_dn), self._p(send_up))

    def migrate_unpack(self, recv_dn, recv_up):
        self.lib.prs_slab_migrate_unpack(self._ref, self._p(recv_dn), self._p(recv_up))

    def sort(self):
        self.lib.prs_slab_sort(self._ref)

    def gather(self):
        self.lib.prs_slab_gather(self._ref)

    def halo_pack(self, send_dn, send_up):
        self.lib.prs_slab_halo_pack(self._ref, self._p(send_dn), self._p(send_up))

    def halo_unpack(self, recv_dn, recv_up):
        self.lib.prs_slab_halo_unpack(self._ref, self._p(recv_dn), self._p(recv_up))

    def cell_table(self):
        self.lib.prs_slab_cell_table(self._ref)

    def collide(self, dt):
        self.lib.prs_slab_collide(self._ref, dt)

    def min_light_distance(self, out):
        self.lib.prs_slab_min_light_distance(self._ref, self._p(out))

    def update_phase(self, spacing, min_d):
        self.lib.prs_slab_update_phase(self._ref, spacing, self._p(min_d))

    def add_noise(self, std):
        self.lib.prs_slab_add_noise(self._ref, std)


class _State:
    pass


class SlabSim:
    """One rank's slab of the swarm.  `backend` supplies the compute calls (CudaBackend or a test
    stand-in); `group` is the torch.distributed process group (None = default)."""

    def __init__(self, params, opt, backend, rank, world, device, pos, gid, rows, capacity=None, halo_cap=None,
                 mig_cap=None, group=None, exchange="nccl", native_step=True, overlap_exchange=True, wrap=False,
                 fused_exchange=True):
        self.p, self.opt, self.be = params, opt, backend
        self.rank, self.world, self.dev, self.group = rank, world, device, group
        # wrap: the slabs form a ring in the row index (the cell hash wraps around the grid, SURVEY.md Q9 — e.g. the reference's
        # own world, +-64 on a 512-row grid of 0.235: robots above y = 56.3 share rows 0.. with the bottom of the world)
        self.wrap = bool(wrap) and world > 1
        self.nb_dn = (rank - 1) % world if (self.wrap or rank > 0) else None
        self.nb_up = (rank + 1) % world if (self.wrap or rank < world - 1) else None
        self.R_lo, self.R_hi = rows[rank], rows[rank + 1]
        self.GX, self.GY = int(params.gridSize.x), int(params.gridSize.y)
        n = int(len(gid))
        cap = self.cap = capacity or int(n * 1.25) + 65536
        self.halo_cap = halo_cap or max(65536, int(cap * 0.1))
        self.mig_cap = mig_cap or max(4096, cap // 128)
        f32, i32 = torch.float32, torch.int32
        s = self.s = _State()
        z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
        s.pos, s.vel = z((cap, 2), f32), z((cap, 2), f32)
        s.rad, s.phase, s.fa, s.fr = z(cap, f32), z(cap, f32), z(cap, f32), z(cap, f32)
        s.dead, s.gid, s.hash, s.scratch = z(cap, i32), z(cap, i32), z(cap, i32), z(cap, i32)
        s.rng = z((cap, 12), i32)
        s.pos[:n] = torch.from_numpy(np.ascontiguousarray(pos)).to(device)
        s.gid[:n] = torch.from_numpy(np.ascontiguousarray(gid).astype(np.int32)).to(device)
        s.rad[:n] = float(np.float32(params.min_radius))
        if int(params.nDead) == -1:
            # object transport: the last robot (global id nCells - 1) is the transported object — larger, never
            # oscillating (reset(), particlebot.cpp:784-791); whichever rank holds it marks it
            obj = s.gid[:n] == int(params.nCells) - 1
            s.rad[:n][obj] = float(np.float32(params.min_radius) * np.float32(params.radFactor))
            s.dead[:n][obj] = 1
        ncat = cap + 2 * self.halo_cap
        self.pr, self.svel, self.hash_cat = z((ncat, 4), f32), z((ncat, 2), f32), z(ncat, i32)
        self.index_sorted = z(cap, i32)
        self.cs = torch.full((int(params.numCells),), -1, dtype=i32, device=device)
        self.ce = z(int(params.numCells), i32)
        self.counts = z(16, i32)
        self.counts[prs.SC_N] = n
        self.lists = z(6 * self.mig_cap, i32)
        self.min_d = z(16, f32)
        mw, hw = 1 + prs.SLAB_MIG_WORDS * self.mig_cap, 1 + prs.SLAB_HALO_WORDS * self.halo_cap
        self.mig_send = [z(mw, i32), z(mw, i32)]      # [down, up]
        self.mig_recv = [z(mw, i32), z(mw, i32)]
        self.halo_send = [z(hw, i32), z(hw, i32)]
        self.halo_recv = [z(hw, i32), z(hw, i32)]
        self.time = np.float32(0.0)
        self.sorted_once = False
        self.be.bind(self)
        self.be.rng_setup(n)
        self.exchange = exchange
        self.seq = {"halo": 0, "mig": 0}
        self.ctx = None
        if exchange == "p2p":
            self._setup_p2p(mw, hw)
        elif exchange != "nccl":
            raise ValueError(exchange)
        if self.exchange == "p2p" and native_step and isinstance(backend, CudaBackend):
            self._setup_native_step(mw, hw, overlap_exchange, fused_exchange)

    # ---- the whole step in the library (prs_slab_step): this class only keeps the process-group plumbing ------------
    def _setup_native_step(self, mw, hw, overlap_exchange, fused_exchange=True):
        c = prs.SlabCtx()
        c.slab = self.be.slab
        c.mailbox, c.peer_dn, c.peer_up = self.mailbox, self.peer[0], self.peer[1]
        c.mw, c.hw = mw, hw
        for i in range(2):
            c.scratch_mig[i] = self.mig_send[i].data_ptr()
            c.scratch_halo[i] = self.halo_send[i].data_ptr()
        c.d_min_d = self.min_d.data_ptr()

        def allreduce_min(dev_ptr, user):
            # the one collective of the path: MIN of a float over the ranks, every phase_update_interval
            dist.all_reduce(self.min_d[:1], op=dist.ReduceOp.MIN, group=self.group)

        self._allreduce_cb = prs.ALLREDUCE_MIN_FN(allreduce_min)     # keep the callback object alive
        c.allreduce_min = self._allreduce_cb
        c.overlap_exchange = 1 if overlap_exchange else 0
        c.fused_exchange = 1 if fused_exchange else 0
        c.time, c.sorted_once = 0.0, 0
        self.ctx = c

    # ---- peer-to-peer mailboxes (exchange="p2p") ---------------------------------------------------
    def _setup_p2p(self, mw, hw):
        """One mailbox per rank, mapped by both neighbours through CUDA IPC.  Layout (uint32 words):
        for source s in (0 = from the lower neighbour, 1 = from the upper): halo[2 parities][hw],
        mig[2 parities][mw]; then the flag words halo_seq[s], mig_seq[s]."""
        lib = self.be.lib
        self._mw, self._hw = mw, hw
        self._region = 2 * hw + 2 * mw
        words = 2 * self._region + 16
        self.mailbox = int(lib.prs_slab_mailbox_alloc(words))
        handle = C.create_string_buffer(int(lib.prs_ipc_handle_size()))
        lib.prs_ipc_export(C.c_void_p(self.mailbox), handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=self.group)
        self.peer = [None, None]     # [lower, upper] neighbour's mailbox in this process's address space
        ok = 1
        opened = {}
        for side, nb in ((0, self.nb_dn), (1, self.nb_up)):
            if nb is not None:
                if nb not in opened:      # a ring of two: both neighbours are the same rank, its mailbox is mapped once
                    opened[nb] = lib.prs_ipc_open(C.create_string_buffer(handles[nb], len(handles[nb])))
                ptr = opened[nb]
                self.peer[side] = int(ptr) if ptr else None
                ok &= int(bool(ptr))
        self._peer_maps = [int(p_) for p_ in opened.values() if p_]
        torch.cuda.synchronize()
        # every rank must use the same exchange: if any mapping failed, all fall back to NCCL
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            for pp in self._peer_maps:
                lib.prs_ipc_close(C.c_void_p(pp))
            lib.prs_slab_mailbox_free(C.c_void_p(self.mailbox))
            self.mailbox, self.peer, self.exchange = None, [None, None], "nccl"
            if self.rank == 0:
                import sys
                print("multigpu: peer-to-peer mapping unavailable, using the NCCL exchange", file=sys.stderr, flush=True)
        dist.barrier(group=self.group)

    def _mb_buf(self, base, src, kind, parity):
        off = src * self._region + (parity * self._hw if kind == "halo" else 2 * self._hw + parity * self._mw)
        return base + 4 * off

    def _mb_flag(self, base, src, kind):
        return base + 4 * (2 * self._region + (0 if kind == "halo" else 2) + src)

    def _p2p_targets(self, kind, send):
        """where this exchange's records go: the neighbour's mailbox (its "from the other side" region),
        or the local scratch buffer when there is no neighbour; advances the sequence number"""
        self.seq[kind] += 1
        q = self.seq[kind]
        par = q & 1
        dn = self._mb_buf(self.peer[0], 1, kind, par) if self.peer[0] else send[0]   # I am the lower rank's UPPER neighbour
        up = self._mb_buf(self.peer[1], 0, kind, par) if self.peer[1] else send[1]
        return q, par, dn, up

    def _p2p_publish_and_wait(self, kind, q):
        be = self.be
        be.signal(self._mb_flag(self.peer[0], 1, kind) if self.peer[0] else None,
                  self._mb_flag(self.peer[1], 0, kind) if self.peer[1] else None, q)
        be.wait(self._mb_flag(self.mailbox, 0, kind) if self.peer[0] else None,
                self._mb_flag(self.mailbox, 1, kind) if self.peer[1] else None, q)

    def close(self):
        if self.ctx is not None:
            self.be.lib.prs_slab_ctx_release(C.byref(self.ctx))
            self.ctx = None
        if self.exchange == "p2p" and getattr(self, "mailbox", None):
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            for pp in self._peer_maps:
                self.be.lib.prs_ipc_close(C.c_void_p(pp))
            dist.barrier(group=self.group)
            self.be.lib.prs_slab_mailbox_free(C.c_void_p(self.mailbox))
            self.mailbox = None

    # ---- helpers ---------------------------------------------------------------------------------
    @staticmethod
    def _gate(time, interval, dt):
        t, T, d = np.float32(time), np.float32(interval), np.float32(dt)
        return bool(t - T * np.floor(t / T) < d)

    def _exchange(self, send, recv):
        """fixed-size buffers to / from the lower (index 0) and upper (index 1) neighbour, one batch;
        ordered on the device stream, the host does not wait for the data"""
        dn, up = self.nb_dn, self.nb_up
        ops = []
        if dn is not None and dn == up:
            # a ring of two: both neighbours are the same rank, messages pair up by order — what I send down arrives as
            # "from above" over there, so the receives are posted in the opposite order of the sends
            ops = [dist.P2POp(dist.isend, send[0], dn, group=self.group), dist.P2POp(dist.irecv, recv[1], dn, group=self.group),
                   dist.P2POp(dist.isend, send[1], dn, group=self.group), dist.P2POp(dist.irecv, recv[0], dn, group=self.group)]
        else:
            if dn is not None:
                ops.append(dist.P2POp(dist.isend, send[0], dn, group=self.group))
                ops.append(dist.P2POp(dist.irecv, recv[0], dn, group=self.group))
            if up is not None:
                ops.append(dist.P2POp(dist.isend, send[1], up, group=self.group))
                ops.append(dist.P2POp(dist.irecv, recv[1], up, group=self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def host_counts(self):
        """the device counters (one small device-to-host copy; NOT part of a step)"""
        return self.counts.cpu().numpy().astype(np.int64)

    @property
    def n(self):
        return int(self.host_counts()[prs.SC_N])

    @property
    def stats(self):
        c = self.host_counts()
        return dict(migrated=int(c[prs.SC_STAT_MIG]), halo=int(c[prs.SC_STAT_HALO]))

    def check(self):
        """raises if a capacity was exceeded or a robot crossed two slabs (sticky device-side flags)"""
        err = int(self.host_counts()[prs.SC_ERR])
        if err:
            raise RuntimeError(f"rank {self.rank}: " + "; ".join(m for bit, m in prs.SLAB_ERRORS.items() if err & bit))

    # ---- dead-cell draw (particlebot.cpp:178-194) and observables ------------------------------------
    def _dead_draw(self):
        """nDead distinct robots, `rand() % remaining` with erase, on the glibc stream seeded with the cfg's
        seed (main.cpp:929) — every rank draws the same GLOBAL ids and marks the ones it owns.  (The
        reference's placement consumes the stream first; swarms set up by a generator start it fresh.)"""
        n_dead, n_total = int(self.p.nDead), int(self.p.nCells)
        if n_dead <= 0:
            return
        libc = C.CDLL(None)
        libc.srand(C.c_uint(int(self.p.seed)))
        alive = list(range(n_total))
        ids = []
        for _ in range(min(n_dead, n_total)):
            ids.append(alive.pop(libc.rand() % len(alive)))
        dead_ids = torch.tensor(ids, dtype=self.s.gid.dtype, device=self.dev)
        self.s.dead[torch.isin(self.s.gid, dead_ids) & (torch.arange(self.cap, device=self.dev) < self.n)] = 1

    def centroid(self):
        """swarm centroid over all ranks (observable; one all_reduce of three doubles, not part of a step)"""
        n = self.n
        acc = torch.zeros(3, dtype=torch.float64, device=self.dev)
        acc[:2] = self.s.pos[:n].double().sum(0)
        acc[2] = n
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        return (acc[:2] / acc[2]).cpu().numpy()

    # ---- one step (Particlebot::update, particlebot.cpp:170-300, cut at the exchanges) ----------------
    def step(self, dt, sort_interval):
        p, be = self.p, self.be
        time = self.time
        if np.float32(time) >= np.float32(p.time_to_dead) and np.float32(time) < np.float32(p.time_to_dead) + np.float32(dt):
            self._dead_draw()
        if self.ctx is not None:
            # the step itself is the library's: K1, migration, sort, gather, halo exchange under the interior collide, edge bands
            err = self.be.lib.prs_slab_step(C.byref(self.ctx), float(dt), float(sort_interval))
            self.time = np.float32(self.ctx.time)
            self.sorted_once = bool(self.ctx.sorted_once)
            if err:
                raise RuntimeError(f"rank {self.rank}: " + "; ".join(msg for bit, msg in prs.SLAB_ERRORS.items() if err & bit))
            return
        phase_step = self._gate(time, p.phase_update_interval, dt)
        sort_step = self._gate(time, sort_interval, dt) or not self.sorted_once
        if phase_step:
            be.min_light_distance(self.min_d)
            dist.all_reduce(self.min_d[:1], op=dist.ReduceOp.MIN, group=self.group)
            be.update_phase(2.0 * float(np.float32(p.min_radius)), self.min_d)
            if p.phase_std:
                be.add_noise(float(p.phase_std))
        be.k1(float(time), float(dt), sort_step)
        p2p = self.exchange == "p2p"
        if sort_step:
            if p2p:
                q, par, dn, up = self._p2p_targets("mig", self.mig_send)
                be.migrate_pack(dn, up)                      # records land in the neighbours' HBM
                self._p2p_publish_and_wait("mig", q)
                be.migrate_unpack(self._mb_buf(self.mailbox, 0, "mig", par), self._mb_buf(self.mailbox, 1, "mig", par))
            else:
                be.migrate_pack(self.mig_send[0], self.mig_send[1])
                self._exchange(self.mig_send, self.mig_recv)
                be.migrate_unpack(self.mig_recv[0], self.mig_recv[1])
            be.sort()
            self.sorted_once = True
        be.gather()
        if p2p:
            q, par, dn, up = self._p2p_targets("halo", self.halo_send)
            be.halo_pack(dn, up)
            self._p2p_publish_and_wait("halo", q)
            be.halo_unpack(self._mb_buf(self.mailbox, 0, "halo", par), self._mb_buf(self.mailbox, 1, "halo", par))
        else:
            be.halo_pack(self.halo_send[0], self.halo_send[1])
            self._exchange(self.halo_send, self.halo_recv)
            be.halo_unpack(self.halo_recv[0], self.halo_recv[1])
        be.cell_table()
        be.collide(float(dt))
        self.time = np.float32(time + np.float32(dt))

    # ---- re-balancing (slab boundaries by equal robot count) ---------------------------------
    _REC = (("pos", 2), ("vel", 2), ("rad", 1), ("phase", 1), ("fa", 1), ("fr", 1), ("dead", 1), ("gid", 1), ("hash", 1), ("rng", 12))

    def rebalance(self):
        """Moves the slab boundaries so that every rank owns about the same number of robots again, and the robots to
        their new owners with their whole state (the migration record: positions, velocities, radius, phase, force sums,
        dead flag, global id, generator state).  A COLLECTIVE control-plane operation between two steps — every rank
        calls it, typically every few thousand steps of a drifting swarm; the data path of a step is unchanged.  The
        histogram of robots per grid row is summed over the ranks, the cuts are placed on its cumulative count
        (`balanced_rows`), robots whose row now belongs to another rank are exchanged, and the next step hashes and sorts
        (ownership is defined by the sorted order).  Results do not depend on the decomposition, so a re-balanced run
        stays bit-equal to the single-GPU run."""
        s, n = self.s, self.n
        y = s.pos[:n, 1].cpu().numpy()
        rows = balanced_rows(self.p, y, self.world, self.group)
        row = np.clip(grid_row_of(y, self.p), 0, self.GY - 1)
        owner = np.searchsorted(np.array(rows[1:-1]), row, "right") if self.world > 1 else np.zeros(n, np.int64)
        rec = torch.cat([getattr(s, name)[:n].reshape(n, w).view(torch.int32) for name, w in self._REC], 1).cpu().numpy()
        parts = {int(d): rec[owner == d] for d in np.unique(owner) if int(d) != self.rank}
        inbox = [None] * self.world
        dist.all_gather_object(inbox, parts, group=self.group)
        arrivals = [p_[self.rank] for r_, p_ in enumerate(inbox) if r_ != self.rank and self.rank in p_]
        new = np.concatenate([rec[owner == self.rank]] + arrivals, 0) if arrivals else rec[owner == self.rank]
        n_new = int(new.shape[0])
        if n_new > self.cap:
            raise RuntimeError(f"rank {self.rank}: {n_new} robots after re-balancing exceed the slab capacity {self.cap}")
        t = torch.from_numpy(np.ascontiguousarray(new)).to(self.dev)
        col = 0
        for name, w in self._REC:
            dst = getattr(s, name)
            dst[:n_new] = t[:, col:col + w].contiguous().view(dst.dtype).reshape((n_new,) + tuple(dst.shape[1:]))
            col += w
        self.counts[prs.SC_N] = n_new
        self.R_lo, self.R_hi = rows[self.rank], rows[self.rank + 1]
        self.be.bind(self)
        self.sorted_once = False                      # the next step hashes, migrates nothing and sorts
        if self.ctx is not None:
            self.ctx.slab = self.be.slab
            self.ctx.sorted_once = 0
        return dict(rows=rows, moved_out=int((owner != self.rank).sum()), n=n_new)

    # ---- steps with the state held by the HOST (what bench.py's e2e measures at N > 1) -------------------
    def time_host_steps(self, dt, sort_interval, steps):
        """`steps` steps in which this rank's pos / vel / rad come from pinned host memory before the step and go back
        to it after the step; returns the wall seconds of the slowest rank (barrier on both sides)"""
        import time as _time
        n_own = self.n
        h = {k: torch.empty((n_own,) + tuple(v.shape[1:]), dtype=v.dtype).pin_memory() for k, v in
             (("pos", self.s.pos), ("vel", self.s.vel), ("rad", self.s.rad))}
        for k, v in h.items():
            v.copy_(getattr(self.s, k)[:n_own])
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        lib = getattr(self.be, "lib", None)
        overlap = self.ctx is not None and lib is not None
        t0 = _time.perf_counter()
        for _ in range(steps):
            if overlap:
                # asynchronous uploads on the launching stream; positions and radii travel back on a second stream as soon
                # as they are final in their slots (after K1 / the migration), under sort, gather and collide; velocities
                # after collide.  Slots [0, n_own) are round-tripped: the device stays the truth across migrations.
                # (the uploads bring back exactly what the last step sent out, so the library's "swarm is sparse" knowledge
                # stays valid: plain stream copies, not prs_h2d_async, which would withdraw the binned sort's admission)
                for k, v in h.items():
                    getattr(self.s, k)[:n_own].copy_(v, non_blocking=True)
                lib.prs_arm_k1_event(1)
                self.step(dt, sort_interval)
                lib.prs_arm_k1_event(0)
                lib.prs_d2h_async(C.c_void_p(h["pos"].data_ptr()), C.c_void_p(self.s.pos.data_ptr()), h["pos"].numel() * 4, 1)
                lib.prs_d2h_async(C.c_void_p(h["rad"].data_ptr()), C.c_void_p(self.s.rad.data_ptr()), h["rad"].numel() * 4, 1)
                lib.prs_d2h_async(C.c_void_p(h["vel"].data_ptr()), C.c_void_p(self.s.vel.data_ptr()), h["vel"].numel() * 4, 0)
                lib.prs_host_step_sync()
            else:
                for k, v in h.items():
                    getattr(self.s, k)[:n_own].copy_(v, non_blocking=True)
                self.step(dt, sort_interval)
                for k, v in h.items():
                    v.copy_(getattr(self.s, k)[:n_own], non_blocking=True)
                torch.cuda.synchronize()
        dist.barrier(group=self.group)
        t = torch.tensor([_time.perf_counter() - t0], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

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


def make_hex_slab(params, opt, geom, backend_factory, rank, world, device, seed, jitter, group=None, exchange="nccl"):
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
    # halo = HALO_ROWS grid rows of ~nx * cell / lattice-row-pitch robots each; 1.6x head room
    per_grid_row = nx * float(params.cellSize.y) / (pitch * 0.8660254)
    halo_cap = int(1.6 * HALO_ROWS * per_grid_row) + 4096
    return SlabSim(params, opt, be, rank, world, device, pos[keep], ids[keep], rows,
                   capacity=int(n_expected * 1.25) + 65536, halo_cap=halo_cap,
                   mig_cap=max(4096, int(0.5 * per_grid_row)), group=group, exchange=exchange)


# --------------------------------------------------------------------------------------------------
# decomposition self-check: a small swarm with migrations on the live ranks against the single-GPU path
SELFCHECK = dict(nx=256, ny=192, pitch=0.17, steps=40, dead=150, light=(-30.0, 0.0), seed=5555)


def selfcheck_config(cfg_path):
    """the 49 152-robot swarm of tests/test_multigpu_gpu.py: example.cfg physics in the reference's +-64 world,
    150 dead robots drawn on step 0, a velocity field that drives robots across the slab boundaries"""
    p, o = prs.load_cfg(cfg_path)
    p.nCells = SELFCHECK["nx"] * SELFCHECK["ny"]
    p.nDead = SELFCHECK["dead"]
    p.light_x, p.light_y = SELFCHECK["light"]
    return p, o, dict(nx=SELFCHECK["nx"], ny=SELFCHECK["ny"], pitch=SELFCHECK["pitch"], half=64.0)


def selfcheck_velocity(gid):
    v = np.zeros((len(gid), 2), np.float32)
    v[:, 1] = (1.5 * np.sin(0.37 * np.asarray(gid).astype(np.float64))).astype(np.float32)
    return v


def selfcheck_vs_single_gpu(cfg_path, rank, world, device, exchange="p2p", steps=None, group=None):
    """Runs the self-check swarm on the `world` live ranks (slab engine) and on rank 0 alone (fused single-GPU path)
    and compares positions, velocities, radii and phases bit for bit.  Rank 0 returns
    {"bit_equal", "robots", "steps", "migrated", "halo_robots"}; the other ranks return None.  Every rank must call it."""
    steps = steps or SELFCHECK["steps"]
    p, o, geom = selfcheck_config(cfg_path)
    n_total = int(p.nCells)
    jitter = 0.01 * p.max_radius
    sim = make_hex_slab(p, o, geom, CudaBackend, rank, world, device, SELFCHECK["seed"], jitter, group=group, exchange=exchange)
    n0 = sim.n
    sim.s.vel[:n0] = torch.from_numpy(selfcheck_velocity(sim.s.gid[:n0].cpu().numpy())).to(device)
    for _ in range(steps):
        sim.step(o.timestep, o.timestep)
    sim.check()
    got = sim.gather_global(n_total)
    st = torch.tensor([sim.stats["migrated"], sim.stats["halo"]], dtype=torch.int64, device=device)
    dist.all_reduce(st, op=dist.ReduceOp.SUM, group=group)
    sim.close()
    out = None
    if rank == 0:
        one = prs.Simulation(p, geom["half"], prs.BACKEND_FUSED)
        one.srand(p.seed)                       # main.cpp:929 — the dead draw continues this stream
        one.init_hex(geom["nx"], geom["ny"], geom["pitch"], jitter, SELFCHECK["seed"])
        one.set(prs.VELOCITY, selfcheck_velocity(np.arange(n_total)))
        for _ in range(steps):
            one.update(o.timestep, o.timestep)
        same = all(np.array_equal(got[k].view(np.uint32), one.get(w).view(np.uint32))
                   for k, w in (("pos", prs.POSITION), ("vel", prs.VELOCITY), ("rad", prs.RADII), ("phase", prs.PHASE)))
        one.close()
        out = {"bit_equal": bool(same and (got["owner"] >= 0).all()), "robots": n_total, "steps": steps,
               "migrated": int(st[0].item()), "halo_robots": int(st[1].item()), "ranks": world,
               "what": "slab engine on the live ranks vs the fused single-GPU path on rank 0: pos, vel, rad, phase bitwise"}
    dist.barrier(group=group)
    return out
