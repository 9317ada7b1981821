#!/usr/bin/env python
"""bench.py — particle-steps/s of the per-timestep particle-robot update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference|ref-cuda]
                    [--robots-log2 L] [--sort-interval S]

A "step" is one Particlebot::update() (controller -> integrate -> hash -> sort -> reorder ->
collide) over the whole synthetic swarm.  Workloads:
  N=1   S1: 2^20 robots on a 1024x1024 hex lattice (pitch 0.17 — 0.155 explodes under the reference's own
        kernels, profiles/workload_stability_r1.txt — jitter 0.01*max_radius, seed 5555),
        world half extent 128, 2048^2 cells of 0.235, light (-90,0), example.cfg physics,
        sort EVERY step (sort_interval = timestep).
  N>1   S2: 2^26 robots (8192x8192), world half extent 768, 8192^2 cells, slab-decomposed over the
        ranks with halo exchange.
Timing: W warm-up steps, then K steps each bracketed by CUDA events on the launching stream with
an L2 flush (256 MiB write) between steps, outside the events; value = robots*K / sum(step times),
max over ranks.  `back_to_back` is the same K steps without flushes (warm L2), for context.
`e2e` runs the same steps through the C-ABI with HOST buffers (prs_sim_update_host): pinned-host -> device copies of
pos/vel/rad before, device -> host copies of pos/vel/rad after every step, inside the timed region (asynchronous; the
downloads of pos and rad overlap sort + collide); `e2e.reference_api` is the same through the reference's blocking calls.
`--impl reference` times the CPU oracle port (OpenMP, all host threads) — the reference ships no
CPU path; `--impl ref-cuda` (extra) times the reference's OWN kernels compiled verbatim
(oracle/_ref) on R1, the largest hex block its hard-coded +-64 world holds (640x640 robots);
`--workload r1` runs this repo's path on the same swarm.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PITCH = 0.17   # BASELINE.md S1 says 0.155 (= 2*min_radius): that crystal is numerically unstable in the reference's DEM model
JITTER_FRAC = 0.01
SEED = 5555
STAGES = ["controller+integrate+hash", "sort", "reorder+celltable", "collide", "phase", "exchange"]
# every GPU workload is timed from this step on, whatever --warmup is (the swarm needs ~250 steps for its contact
# network to form: collide is cheaper before that); the warm-up steps are the last W steps before it
EVOLVE_TO = 260


# ------------------------------------------------------------------------------------------------
class _Run:
    """timestep / sort_interval of the cfg (stand-in for prs.RunOptions on the arm that must not load the product library)"""

    def __init__(self, d):
        self.timestep, self.sort_interval = float(np.float32(d["timestep"])), float(np.float32(d["sort_interval"]))


def swarm_config(prs, log2n, world64=False, nx=None, ny=None, pitch=None):
    """SimParams + hex-block geometry of the synthetic swarm with 2^log2n robots (or nx*ny).
    prs = the product package, or None: parameters from oracle/params.py (pure Python; the reference arm)."""
    cfg = os.path.join(ROOT, "examples", "example.cfg")
    if prs is None:
        from oracle import params as op
        p, run = op.load_cfg(cfg)
        o = _Run(run)
    else:
        p, o = prs.load_cfg(cfg)
    pitch = PITCH if pitch is None else pitch
    if nx is None:
        nx = 1 << ((log2n + 1) // 2)
        ny = 1 << (log2n // 2)
    p.nCells = nx * ny
    w, h = nx * pitch, ny * pitch * 0.8660254
    if world64:
        half, grid = 64.0, 512          # the reference's hard-coded world (kernel_impl.cuh:75-97, main.cpp:937)
        light = (-60.0, 0.0)
    else:
        half = float(np.ceil(max(w, h) / 2 * 1.2 / 64.0) * 64.0)
        grid = 1 << int(np.ceil(np.log2(2 * half / 0.235)))
        light = {20: (-90.0, 0.0), 26: (-700.0, 0.0)}.get(log2n, (-0.7 * half, 0.0))
    if prs is None:
        op.set_world(p, grid, half)
    else:
        prs.lib().prs_params_set_world(C.byref(p), grid, half)
    p.light_x, p.light_y = light
    p.max_time = 1e30
    name = {20: "S1", 26: "S2"}.get(log2n, f"S(2^{log2n})")
    return p, o, dict(name=name, nx=nx, ny=ny, half=half, grid=grid, light=light, pitch=pitch)


def workload_text(geom, n):
    """the `config.workload` string, shared by both arms"""
    return (f"{geom['name']}: {n} robots, hex {geom['nx']}x{geom['ny']} pitch {geom['pitch']}, "
            f"world +-{geom['half']:g}, grid {geom['grid']}^2, light {geom['light']}")


def algorithmic_bytes(p, sort_every_step=True):
    """B_alg = 158 + 16*P + 4*C/N per particle-step (146 + 4*C/N without the sort)."""
    bits = int(np.ceil(np.log2(p.numCells)))
    passes = (bits + 7) // 8
    per = {"controller+integrate+hash": 60 if sort_every_step else 52, "sort": 4 + 16 * passes if sort_every_step else 0,
           "reorder+celltable": 51 + 4.0 * p.numCells / p.nCells, "collide": 43, "phase": 0, "exchange": 0}
    return per, sum(per.values()), passes


class ClockSampler:
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md's clocks line), polled
    through NVML every 5 ms from a thread (nvidia-smi as a fallback: it needs a second or more to come
    up on an 8-GPU box, longer than a short timed region).  `mark()` brackets the timed region; the
    report uses the samples inside it (all samples taken under load if the region was too short)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []          # (time, sm_mhz, max_mhz, reasons bitmask)
        self.t0 = self.t1 = None
        self._stop = False
        self.thread = None
        self.proc = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].strip().isdigit() else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)

            def poll():
                while not self._stop:
                    try:
                        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        try:
                            rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.samples.append((time.time(), float(sm), float(mx), int(rs)))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "nvml"
        except Exception:
            self._start_smi()

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.source = "nvidia-smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)

            def read():
                bits = (0x8, 0x40, 0x20, 0x4)
                for line in self.proc.stdout:
                    f = [x.strip() for x in line.split(",")]
                    try:
                        rs = sum(b for b, v in zip(bits, f[2:6]) if v.lower().startswith("active"))
                        self.samples.append((time.time(), float(f[0]), float(f[1]), rs))
                    except (ValueError, IndexError):
                        pass

            self.thread = threading.Thread(target=read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        """call at the start and at the end of the timed region"""
        if self.t0 is None:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def stop(self):
        self._stop = True
        if self.proc:
            self.proc.terminate()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"], "samples": 0}
        inside = [x for x in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= x[0] <= self.t1 + 0.005]
        scope = "timed region"
        if len(inside) < 2:
            inside, scope = self.samples, "warm-up + timed region + stage timing (timed region shorter than two polls)"
        mask = 0
        for x in inside:
            mask |= x[3]
        return {"sm_mhz": float(np.median([x[1] for x in inside])), "sm_max_mhz": max(x[2] for x in inside),
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(inside), "scope": scope,
                "source": self.source}


def ncu_traffic(kernel):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json)"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(kernel)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
def cpu_oracle_throughput(p, o, geom, pos0, steps, threads, warm=1, min_seconds=0.0):
    """particle-steps/s of the CPU oracle port on this swarm (test infrastructure used as baseline);
    runs `steps` steps, then keeps stepping until min_seconds of CPU work have been timed"""
    from oracle import binding as ob
    L = ob.lib()
    L.prso_set_threads(threads)
    s = ob.OracleSim(p, geom["half"])
    s.view("pos")[:] = pos0
    s.view("rad")[:] = p.min_radius
    for _ in range(warm):
        s.update(o.timestep, o.timestep)
    t0 = time.perf_counter()
    done = 0
    while done < steps or time.perf_counter() - t0 < min_seconds:
        s.update(o.timestep, o.timestep)
        done += 1
    dt = time.perf_counter() - t0
    s.close()
    L.prso_set_threads(1)
    return p.nCells * done / dt, dt, done


def hex_positions(p, geom):
    """Same generator as Particlebot::initHexBlock, through the library (no GPU needed for the maths
    but the object needs one; used by the GPU arms) — CPU copy for the reference arm."""
    nx, ny = geom["nx"], geom["ny"]
    n = nx * ny
    i = np.arange(n, dtype=np.uint64)
    ix, iy = (i % nx).astype(np.float32), (i // nx)
    row = np.float32(geom["pitch"] * 0.8660254037844386)
    pitch = np.float32(geom["pitch"])
    x0 = np.float32(-0.5) * (np.float32(nx - 1) * pitch + np.float32(0.5) * pitch)
    y0 = np.float32(-0.5) * np.float32(ny - 1) * row

    def mix(z):
        z = (z + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))

    with np.errstate(over="ignore"):
        h = mix((np.uint64(SEED) << np.uint64(32)) ^ i)
    jit = np.float32(JITTER_FRAC * p.max_radius)
    jx = ((h & np.uint64(0xFFFFFF)).astype(np.float32) / np.float32(8388608.0) - np.float32(1.0)) * jit
    jy = (((h >> np.uint64(24)) & np.uint64(0xFFFFFF)).astype(np.float32) / np.float32(8388608.0) - np.float32(1.0)) * jit
    x = x0 + ix * pitch + np.where((iy & 1) == 1, np.float32(0.5) * pitch, np.float32(0.0)).astype(np.float32) + jx
    y = y0 + iy.astype(np.float32) * row + jy
    return np.stack([x, y], 1).astype(np.float32)



def timed_steps(torch, sim, dt, sort_interval, steps, warmup, flush, stat="median"):
    """ms per step of sim.update over `steps` steps (CUDA events on the current stream, L2 flushed between steps)"""
    stream = torch.cuda.current_stream()
    for _ in range(warmup):
        sim.update(dt, sort_interval)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record(stream)
        sim.update(dt, sort_interval)
        b.record(stream)
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in ev]
    return float(np.median(t) if stat == "median" else np.mean(t))   # median: the reference's per-sort cudaMalloc/cudaFree makes outliers


def pairs_per_robot(prs, sim, p):
    """ordered neighbour pairs per robot that collide evaluates (5x5 stencil, wrapped), from the current hashes"""
    n = int(p.nCells)
    h = sim.get(prs.HASH).astype(np.int64)
    gdim = int(p.gridSize.x)
    cnt = np.bincount(h, minlength=gdim * gdim).reshape(gdim, gdim).astype(np.int64)
    rows = np.flatnonzero(cnt.sum(1))
    cols = np.flatnonzero(cnt.sum(0))
    if rows.size and rows[0] >= 2 and rows[-1] < gdim - 2 and cols[0] >= 2 and cols[-1] < gdim - 2:
        cnt = cnt[rows[0] - 2:rows[-1] + 3, cols[0] - 2:cols[-1] + 3]     # nothing wraps: work on the occupied box only
    box = sum(np.roll(np.roll(cnt, dy, 0), dx, 1) for dy in range(-2, 3) for dx in range(-2, 3))
    return int((cnt * box).sum() - n)


def one_gpu_swarm_block(torch, prs, flush, log2n, steps, evolve_to=None, pitch=None, warmup=3):
    """a whole synthetic swarm on THIS GPU through the fused path: evolved (untimed) to step `evolve_to`, then `steps`
    steps timed like the primary (events per step, L2 flushed between them, mean)"""
    evolve_to = EVOLVE_TO if evolve_to is None else evolve_to
    p, o, geom = swarm_config(prs, log2n, pitch=pitch)
    n = int(p.nCells)
    sim = prs.Simulation(p, geom["half"], prs.BACKEND_FUSED)
    sim.init_hex(geom["nx"], geom["ny"], geom["pitch"], JITTER_FRAC * p.max_radius, SEED)
    for k in range(max(evolve_to - warmup, 0)):
        sim.update(o.timestep, o.timestep)
        if k == 3:
            sim.sync()
    ms = timed_steps(torch, sim, o.timestep, o.timestep, steps, min(warmup, evolve_to), flush, stat="mean")
    pairs = pairs_per_robot(prs, sim, p) / n
    finite = bool(np.isfinite(sim.get(prs.POSITION)).all())
    sim.close()
    _, b_alg, passes = algorithmic_bytes(p, True)
    peak, _ = measured_peak()
    return {"workload": workload_text(geom, n), "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "particle-steps/s",
            "steps": steps, "timed_state_step": evolve_to, "state": f"steps {evolve_to}..{evolve_to + steps} of the swarm started from the lattice",
            "ordered_pairs_per_robot": pairs, "state_finite": finite,
            "roofline_step": {"alg_bytes_per_particle_step": b_alg, "achieved": b_alg * n / (ms * 1e-3) / 1e9, "peak": peak,
                              "unit": "GB/s", "frac": b_alg * n / (ms * 1e-3) / 1e9 / peak}}


def ref_cuda_block(torch, prs, flush, steps=30, warmup=5):
    """The reference's OWN kernels (oracle/_ref, compiled verbatim for sm_100a) beside this repo's path on
    R1 — the largest hex block the reference's hard-coded +-64 world holds (640x640 robots)."""
    from oracle import binding as ob
    if not os.path.exists(ob.REFCUDA_PATH):
        return {"unavailable": "oracle/_ref/libprs_refcuda.so not built"}
    out = {"workload": "R1: 409600 robots, hex 640x640 pitch 0.17, reference world +-64, grid 512^2, sort every step"}
    for name, backend, ext in (("reference_kernels", prs.BACKEND_EXTERNAL, ob.REFCUDA_PATH), ("this_repo", prs.BACKEND_FUSED, None)):
        p, o, geom = swarm_config(prs, 0, world64=True, nx=640, ny=640)
        sim = prs.Simulation(p, geom["half"], backend, ext)
        sim.init_hex(geom["nx"], geom["ny"], geom["pitch"], JITTER_FRAC * p.max_radius, SEED)
        ms = timed_steps(torch, sim, o.timestep, o.timestep, steps, warmup, flush)
        sim.close()
        out[name] = {"ms_per_step": ms, "value": p.nCells / (ms * 1e-3), "unit": "particle-steps/s"}
    out["speedup"] = out["reference_kernels"]["ms_per_step"] / out["this_repo"]["ms_per_step"]
    return out


# ------------------------------------------------------------------------------------------------
def run_reference_cpu(args):
    """--impl reference: the reference has no CPU implementation; this times the oracle port with
    every host thread on a bounded sample of the same workload (same lattice, density, physics)."""
    from oracle import binding as ob
    prs = None      # this arm must not map the product library: parameters come from oracle/params.py
    threads = len(os.sched_getaffinity(0))
    budget_s = 120.0
    log2n = args.robots_log2 or (20 if args.gpus == 1 else 26)
    # probe at 2^14 robots, then take the largest power of four <= the workload that fits the budget
    p, o, geom = swarm_config(prs, 14)
    rate, _, _ = cpu_oracle_throughput(p, o, geom, hex_positions(p, geom), 3, threads)
    sample = 14
    while sample + 2 <= min(log2n, 22) and (1 << (sample + 2)) * (args.steps + args.warmup) / rate < budget_s:
        sample += 2            # at most 2^22 robots: bounded host memory and time whatever K and W are
    p, o, geom = swarm_config(prs, sample)
    pos0 = hex_positions(p, geom)
    L = ob.lib()
    L.prso_set_threads(threads)
    s = ob.OracleSim(p, geom["half"])
    s.view("pos")[:] = pos0
    s.view("rad")[:] = p.min_radius
    for _ in range(args.warmup):
        s.update(o.timestep, o.timestep)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.update(o.timestep, o.timestep)
    dt = time.perf_counter() - t0
    value = p.nCells * args.steps / dt
    sample_txt = f"{geom['name']} lattice cut to 2^{sample} robots ({geom['nx']}x{geom['ny']}), {args.steps} steps, sort every step"
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(swarm_config(prs, log2n)[2], 1 << log2n),
                   "sort_interval": "timestep (sort every step)", "cpu_sample": f"2^{sample} robots of the same lattice"},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference ships no CPU path: OpenMP C++ restatement (oracle/prs_oracle.cpp) on the host cores",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import particlerobotsimulations_b200 as prs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        return run_slabs(args, rank, world, local_rank)

    torch.cuda.set_device(local_rank)
    lib = prs.lib()
    ref_cuda = args.impl == "ref-cuda"
    log2n = args.robots_log2 or 20
    r1 = ref_cuda or args.workload == "r1"
    if r1:
        # R1: the largest hex block the reference's hard-coded +-64 world / 512^2 grid holds
        p, o, geom = swarm_config(prs, 0, world64=True, nx=640, ny=640)
        geom["name"] = "R1"
    else:
        p, o, geom = swarm_config(prs, log2n)
    n = int(p.nCells)
    sort_interval = o.timestep if args.sort_interval is None else args.sort_interval
    stream = torch.cuda.current_stream()
    lib.prs_set_stream(C.c_void_p(stream.cuda_stream))
    if args.collide_tile is not None:
        lib.prs_set_collide_tile(args.collide_tile)
    if args.pdl is not None:
        lib.prs_set_pdl(args.pdl)

    if ref_cuda:
        from oracle import binding as ob
        if not os.path.exists(ob.REFCUDA_PATH):
            print(json.dumps({"impl": "ref-cuda", "unavailable": "oracle/_ref/libprs_refcuda.so not built"}))
            return
        sim = prs.Simulation(p, geom["half"], prs.BACKEND_EXTERNAL, ob.REFCUDA_PATH)
    else:
        sim = prs.Simulation(p, geom["half"], prs.BACKEND_FUSED)
    sim.init_hex(geom["nx"], geom["ny"], geom["pitch"], JITTER_FRAC * p.max_radius, SEED)
    pos0 = sim.get(prs.POSITION)
    if args.scramble:   # robot index unrelated to position (the reference's aggregation placement is like that)
        pos0 = pos0[np.random.default_rng(SEED).permutation(n)]
        sim.set(prs.POSITION, pos0)
        geom["name"] += " (robot order scrambled)"

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        sim.update(o.timestep, sort_interval)

    sampler = ClockSampler(local_rank)
    sampler.start()
    evolve_to = max(args.evolve_to, args.warmup)
    for k in range(evolve_to - args.warmup):     # untimed: brings the swarm to the fixed state the timing starts from
        step()
        if k == 3:
            torch.cuda.synchronize()             # the density report arrives: cell binning from here on
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, L2 flushed between steps (outside the event pairs) ----
    sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    lib.prs_launch_count(1)
    torch.cuda.synchronize()
    for a, b in ev:
        flush.fill_(1)
        a.record(stream)
        step()
        b.record(stream)
    torch.cuda.synchronize()
    sampler.mark()
    launches = int(lib.prs_launch_count(0))
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(np.sum(step_ms))
    # back-to-back (warm L2), one event pair around K steps
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(args.steps):
        step()
    b.record(stream)
    torch.cuda.synchronize()
    b2b_ms = a.elapsed_time(b)

    value = n * args.steps / (total_ms * 1e-3)

    # ---- per-stage event timing (native fused path only) ----
    stages, roofline, roofline_step = None, None, None
    peak, peak_src = measured_peak()
    per_bytes, b_alg, passes = algorithmic_bytes(p, sort_interval <= o.timestep)
    if not ref_cuda:
        lib.prs_stage_timing(1)
        for _ in range(args.steps):
            flush.fill_(1)
            step()
        ms = (C.c_float * 6)()
        cnt = (C.c_uint * 6)()
        lib.prs_stage_times(ms, cnt)
        lib.prs_stage_timing(0)
        stages = {}
        for i, name in enumerate(STAGES):
            if cnt[i]:
                avg_us = 1e3 * ms[i] / cnt[i]
                gbs = per_bytes[name] * n / (avg_us * 1e-6) / 1e9 if per_bytes[name] else None
                stages[name] = {"avg_us": avg_us, "alg_bytes_per_robot": per_bytes[name], "achieved_GBps": gbs,
                                "frac_of_hbm_peak": (gbs / peak) if gbs else None}
        dom = max(stages, key=lambda k: stages[k]["avg_us"])
        roofline = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
                    "frac": stages[dom]["frac_of_hbm_peak"], "traffic": ncu_traffic(dom) if log2n == 20 and not r1 else None,
                    "peak_source": peak_src,
                    "note": "collide is FP32/MUFU-issue-bound (IEEE div/sqrt per neighbour pair), not HBM-bound; "
                            "see roofline_step for the whole-step HBM fraction"}
    clocks = sampler.stop()
    step_gbs = b_alg * value / 1e9
    roofline_step = {"bound": "hbm", "alg_bytes_per_particle_step": b_alg, "radix_passes": passes, "achieved": step_gbs,
                     "peak": peak, "unit": "GB/s", "frac": step_gbs / peak, "peak_source": peak_src}

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region ----
    h_pos = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    h_vel = torch.empty((n, 2), dtype=torch.float32).pin_memory()
    h_rad = torch.empty((n,), dtype=torch.float32).pin_memory()
    for which, t in ((prs.POSITION, h_pos), (prs.VELOCITY, h_vel), (prs.RADII, h_rad)):
        lib.prs_sim_get(sim._h, which, t.data_ptr(), t.numel() * 4)
    e2e_steps = max(3, min(args.steps, 50))
    # (1) the host-buffer step of this library: prs_sim_update_host (asynchronous copies; positions and radii
    #     travel back under the sort and collide kernels)
    for _ in range(2):
        lib.prs_sim_update_host(sim._h, h_pos.data_ptr(), h_vel.data_ptr(), h_rad.data_ptr(), h_pos.data_ptr(), h_vel.data_ptr(),
                                h_rad.data_ptr(), o.timestep, sort_interval)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        lib.prs_sim_update_host(sim._h, h_pos.data_ptr(), h_vel.data_ptr(), h_rad.data_ptr(), h_pos.data_ptr(), h_vel.data_ptr(),
                                h_rad.data_ptr(), o.timestep, sort_interval)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": n * e2e_steps / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 20 * n,
           "d2h_bytes_per_step": 20 * n, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "what": "prs_sim_update_host: pos,vel,rad from pinned host -> device, Particlebot::update, pos,vel,rad -> pinned host "
                   "(every step; the downloads of pos and rad overlap sort+collide)"}
    # (2) the same through the reference's own blocking calls (copyArrayToDevice x3, update, copyArrayFromDevice x3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for which, t in ((prs.POSITION, h_pos), (prs.VELOCITY, h_vel), (prs.RADII, h_rad)):
            lib.prs_sim_set(sim._h, which, t.data_ptr(), 0, t.numel() * 4)
        step()
        for which, t in ((prs.POSITION, h_pos), (prs.VELOCITY, h_vel), (prs.RADII, h_rad)):
            lib.prs_sim_get(sim._h, which, t.data_ptr(), t.numel() * 4)
    torch.cuda.synchronize()
    ref_api_s = time.perf_counter() - t0
    e2e["reference_api"] = {"value": n * e2e_steps / ref_api_s, "ms_per_step": 1e3 * ref_api_s / e2e_steps,
                            "what": "copyArrayToDevice(pos,vel,rad) -> Particlebot::update -> copyArrayFromDevice(pos,vel,rad), blocking copies"}

    finite = bool(np.isfinite(sim.get(prs.POSITION)).all())
    # ---- the dominant kernel against the unit that actually binds it (informational, next to the HBM roofline) ----
    roofline_compute = None
    if not ref_cuda and stages and "collide" in stages:
        pairs = pairs_per_robot(prs, sim, p)                   # ordered neighbour pairs of the reference's loop per step
        sm_clock_hz = 1e6 * (clocks.get("sm_mhz") or 1965.0)
        xu_peak = 148 * 16 * sm_clock_hz / 4.0                 # 16 MUFU lanes per SM and clock, 4 MUFU ops per far pair
        rate = pairs / (stages["collide"]["avg_us"] * 1e-6)
        roofline_compute = {"kernel": "collide", "bound": "xu (MUFU) pipe", "ordered_pairs_per_step": pairs,
                            "unordered_pairs_per_step": pairs // 2, "achieved": rate / 2, "peak": xu_peak,
                            "unit": "unordered neighbour pairs/s", "frac": rate / 2 / xu_peak, "frac_ordered_pairs": rate / xu_peak,
                            "what": "floor = every UNORDERED pair evaluated once (F_ji = -F_ij bit for bit) at 4 MUFU ops "
                                    "(rsqrt, lg2, ex2, rcp) per far pair; peak = 148 SMs x 16 MUFU lanes x SM clock / 4. "
                                    "frac_ordered_pairs counts the reference's loop (each pair from both sides)"}
    sim.close()

    # ---- CPU baseline beside it (rank 0, N=1): oracle port, all threads, a few steps of the same swarm ----
    cpu = None
    if not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        cpu_log2 = min(log2n, 20)
        pc, oc, gc = swarm_config(prs, cpu_log2)
        rate, dt, done = cpu_oracle_throughput(pc, oc, gc, hex_positions(pc, gc), 4, threads, min_seconds=12.0)
        cpu = {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port",
               "sample": f"same lattice at 2^{cpu_log2} robots, {done} steps after 1 warm-up, sort every step ({dt:.1f} s of "
                         f"{threads}-thread OpenMP work; oracle/prs_oracle.cpp)"}

    # ---- secondary: the reference's cadence — hash + sort only every 180 s of simulated time ----
    secondary = None
    if not ref_cuda and sort_interval <= o.timestep:
        sim2 = prs.Simulation(p, geom["half"], prs.BACKEND_FUSED)
        sim2.init_hex(geom["nx"], geom["ny"], geom["pitch"], JITTER_FRAC * p.max_radius, SEED)
        if args.scramble:
            sim2.set(prs.POSITION, pos0)
        for k in range(max(evolve_to - 10, 0)):      # same timed state as the primary (the first step sorts, the rest reuse it)
            sim2.update(o.timestep, 180.0)
        ms2 = timed_steps(torch, sim2, o.timestep, 180.0, min(args.steps, 100), 10, flush)
        secondary = {"sort_interval": 180.0, "ms_per_step": ms2, "value": n / (ms2 * 1e-3), "unit": "particle-steps/s",
                     "what": "same swarm, reference cadence: controller+integrate, gather, collide on the step-0 ordering (Q1)"}
        if args.try_fused_gather:   # experiment: K1 + gather as one kernel also at this size
            lib.prs_set_fuse_gather_max(1 << 30)
            ms3 = timed_steps(torch, sim2, o.timestep, 180.0, min(args.steps, 100), 10, flush)
            lib.prs_set_fuse_gather_max(65536)
            secondary["ms_per_step_fused_gather"] = ms3
        sim2.close()

    ref_cuda_cmp = None
    if not ref_cuda and not args.no_ref_cuda:
        ref_cuda_cmp = ref_cuda_block(torch, prs, flush)

    # ---- BASELINE.md's S1 as specified (pitch 0.155 = 2*min_radius), over the steps before the reference's own kernels
    #      blow that crystal up (profiles/workload_stability_r1.txt): steps 10..50 ----
    spec_pitch = None
    if not ref_cuda and not r1 and log2n == 20 and not args.no_extras:
        spec_pitch = one_gpu_swarm_block(torch, prs, flush, 20, 40, evolve_to=10, pitch=0.155)
        spec_pitch["what"] = ("S1 at BASELINE.md's pitch 0.155; numerically unstable under the reference's explicit-Euler DEM "
                              "(NaN by step 300 with the reference's own kernels), hence timed over steps 10..50 only")
    # ---- S2 (2^26 robots) on this one GPU: the same-workload baseline of the multi-GPU lines ----
    s2_one = None
    if not ref_cuda and not r1 and log2n == 20 and not args.no_extras and not args.no_s2:
        s2_one = one_gpu_swarm_block(torch, prs, flush, 26, min(args.steps, 25))

    line = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(geom, n),
                   "sort_interval": "timestep (sort every step)" if sort_interval <= o.timestep else sort_interval,
                   "collide_mode": "exact", "l2": "flushed between timed steps (256 MiB write)",
                   "collide_neighbours": "shared-memory windows staged by TMA bulk copies" if lib.prs_get_collide_tile() else "L1/L2",
                   "programmatic_dependent_launch": bool(lib.prs_get_pdl())},
        "back_to_back": {"value": n * args.steps / (b2b_ms * 1e-3), "ms_per_step": b2b_ms / args.steps, "l2": "warm"},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_step": roofline_step, "roofline_compute": roofline_compute,
        "stages": stages, "cpu_baseline": cpu, "ref_cuda": ref_cuda_cmp, "secondary": secondary, "state_finite": finite,
        "timed_state_step": evolve_to, "secondary_spec_pitch_0155": spec_pitch, "s2_one_gpu": s2_one,
        "scaling_note": "--gpus 1 times S1 (2^20 robots, the single-GPU roofline workload); --gpus N>1 times S2 (2^26 robots, fixed "
                        "total: strong scaling) and carries its own same-workload 1-GPU point (one_gpu_same_workload); "
                        "s2_one_gpu here is that point measured in this run",
    }
    if ref_cuda:
        line["impl"] = "ref-cuda"
        line["config"]["workload"] += " [reference kernels compiled verbatim; reference world +-64, grid 512^2]"
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_sort_only(args):
    """--sort-only: the hand-written onesweep sort (prs_sort_pairs) beside cub::DeviceRadixSort::SortPairs (cross-check
    build, oracle/_build/libprs_cubsort.so — what the reference reaches through thrust::sort_by_key,
    particlebot_cuda.cu:377-382) on the cell keys of the S1 / S2 lattices at 2^20, 2^23 and 2^26 pairs: identical
    output, median of `--steps` runs each (CUDA events, L2 flushed between runs), algorithmic bytes 4 + 16 P per pair."""
    import torch
    import particlerobotsimulations_b200 as prs
    from oracle import binding as ob
    torch.cuda.set_device(0)
    lib = prs.lib()
    stream = torch.cuda.current_stream()
    lib.prs_set_stream(C.c_void_p(stream.cuda_stream))
    cub_path = os.path.join(os.path.dirname(ob.LIB_PATH), "libprs_cubsort.so")
    cub = C.CDLL(cub_path) if os.path.exists(cub_path) else None
    if cub:
        cub.prs_cub_sort_temp_bytes.restype = C.c_size_t
        cub.prs_cub_sort_temp_bytes.argtypes = [C.c_uint, C.c_int, C.c_int]
        cub.prs_cub_sort_pairs.restype = C.c_int
        cub.prs_cub_sort_pairs.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 4 + [C.c_uint, C.c_int, C.c_int, C.c_void_p]
    peak, peak_src = measured_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    reps = max(5, min(args.steps, 30))
    rows = []
    for log2n in (0, 20, 23, 26):
        # 0: R1, 640^2 robots in the reference's own +-64 world (512^2 cells: 18-bit keys, the size every shipped cfg sorts at)
        p, o, geom = swarm_config(prs, 0, world64=True, nx=640, ny=640) if log2n == 0 else swarm_config(prs, log2n)
        if log2n == 0:
            geom["name"] = "R1"
        n = int(p.nCells)
        pos = torch.from_numpy(hex_positions(p, geom)).cuda()
        # cell keys of the lattice (calcHash through the C-ABI), robots in index order: the distribution the step sorts
        keys = torch.empty(n, dtype=torch.int32, device="cuda")
        vals = torch.empty(n, dtype=torch.int32, device="cuda")
        lib.setParameters(C.byref(p))
        lib.calcHash(keys.data_ptr(), vals.data_ptr(), pos.data_ptr(), n)
        bits = int(np.ceil(np.log2(p.numCells)))
        plan = (C.c_int * 6)()
        passes = int(lib.prs_sort_plan(bits, n, plan))           # digits of this sort: 8 bits, or 9 where that saves a pass
        ok, ov = torch.empty_like(keys), torch.empty_like(vals)
        ck, cv = torch.empty_like(keys), torch.empty_like(vals)

        def timed(fn):
            ts = []
            for _ in range(reps):
                flush.fill_(1)
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                fn()
                b_.record(stream)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_))
            return float(np.median(ts))

        own = lambda: lib.prs_sort_pairs(keys.data_ptr(), vals.data_ptr(), ok.data_ptr(), ov.data_ptr(), n, bits)
        own()
        t_own = timed(own)
        row = {"pairs": n, "workload": workload_text(geom, n), "key_bits": bits, "radix_passes": passes,
               "digit_bits": [int(plan[i]) for i in range(passes)], "pairs_per_thread": int(plan[4]),
               "alg_bytes_per_pair": 4 + 16 * passes,
               "onesweep": {"ms": t_own, "GBps": (4 + 16 * passes) * n / (t_own * 1e-3) / 1e9,
                            "frac_of_hbm_peak": (4 + 16 * passes) * n / (t_own * 1e-3) / 1e9 / peak}}
        if cub:
            for name, end_bit in (("cub_key_bits", bits), ("cub_32_bits", 32)):
                tb = cub.prs_cub_sort_temp_bytes(n, 0, end_bit)
                temp = torch.empty(max(tb, 16), dtype=torch.uint8, device="cuda")
                f = lambda: cub.prs_cub_sort_pairs(temp.data_ptr(), tb, keys.data_ptr(), ck.data_ptr(), vals.data_ptr(), cv.data_ptr(),
                                                   n, 0, end_bit, C.c_void_p(stream.cuda_stream))
                assert f() == 0
                t = timed(f)
                pp = (end_bit + 7) // 8
                row[name] = {"ms": t, "end_bit": end_bit, "GBps": (4 + 16 * pp) * n / (t * 1e-3) / 1e9,
                             "frac_of_hbm_peak": (4 + 16 * pp) * n / (t * 1e-3) / 1e9 / peak}
            torch.cuda.synchronize()
            row["identical_output"] = bool(torch.equal(ok, ck) and torch.equal(ov, cv))
            row["onesweep_vs_cub_key_bits"] = row["cub_key_bits"]["ms"] / t_own
            row["onesweep_vs_cub_32_bits"] = row["cub_32_bits"]["ms"] / t_own
        rows.append(row)
        del pos, keys, vals, ok, ov, ck, cv
    print(json.dumps({"metric": "sort: pairs/s (informational, --sort-only)", "impl": "sort-only", "peak_GBps": peak, "peak_source": peak_src,
                      "what": "prs_sort_pairs (hand-written onesweep, csrc/prs_onesweep.cuh) vs cub::DeviceRadixSort::SortPairs on the "
                              "cell keys of the synthetic lattices; cub_32_bits is what thrust::sort_by_key runs (full key width), "
                              "cub_key_bits is CUB told the real key width", "rows": rows}))


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(gpu_index):
    """Pins this process to the CPUs NVML lists as local to its GPU, so that the pinned host buffers of the e2e block are
    first-touched on the GPU's own NUMA node (torchrun binds nothing: with 8 ranks on two sockets half of the host <-> device
    traffic otherwise crosses the socket link).  Returns the number of CPUs bound to, or 0 if nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_slabs(args, rank, world, local_rank):
    """--gpus N (N > 1): S2 (2^26 robots, or --robots-log2) slab-decomposed over the ranks.  Before timing, the slab
    engine is checked bit for bit against the single-GPU path on the live ranks (multigpu.selfcheck_vs_single_gpu);
    after it, rank 0 alone runs the SAME swarm on its one GPU at the same timed state (one_gpu_same_workload)."""
    import torch
    import torch.distributed as dist
    import particlerobotsimulations_b200 as prs
    from particlerobotsimulations_b200 import multigpu

    bound_cpus = bind_to_gpu_numa(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    lib = prs.lib()
    parity = None
    if not args.no_parity_check:
        parity = multigpu.selfcheck_vs_single_gpu(os.path.join(ROOT, "examples", "example.cfg"), rank, world, dev,
                                                  exchange=args.exchange)
        ok = torch.tensor([1 if (parity is None or parity["bit_equal"]) else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if rank == 0:
                print(json.dumps({"metric": "particle-steps/sec", "n_gpus": world, "error": "slab engine differs from the single-GPU path",
                                  "parity_check": parity}))
            dist.destroy_process_group()
            sys.exit(3)
    log2n = args.robots_log2 or 26
    p, o, geom = swarm_config(prs, log2n)
    n_total = int(p.nCells)
    sim = multigpu.make_hex_slab(p, o, geom, multigpu.CudaBackend, rank, world, dev, SEED, JITTER_FRAC * p.max_radius,
                                 exchange=args.exchange)
    sort_interval = o.timestep if args.sort_interval is None else args.sort_interval
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    evolve_to = max(args.evolve_to, args.warmup)
    for _ in range(evolve_to):                 # untimed pre-evolution to the fixed timed state; its last W steps are the warm-up
        sim.step(o.timestep, sort_interval)
    torch.cuda.synchronize()
    dist.barrier()
    sampler.mark()
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
    sampler.mark()
    launches = int(lib.prs_launch_count(0))
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    sim.check()
    st = sim.stats
    n_own = sim.n
    own = torch.tensor([n_own, st["migrated"], st["halo"]], dtype=torch.int64, device=dev)
    owns = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(owns, own)
    steps_run = evolve_to + args.steps

    # ---- per-stage event timing on this rank (kernels only; the neighbour transfers sit between the stages) ----
    lib.prs_stage_timing(1)
    for _ in range(min(args.steps, 10)):
        flush.fill_(1)
        sim.step(o.timestep, sort_interval)
    ms = (C.c_float * 6)()
    cnt = (C.c_uint * 6)()
    lib.prs_stage_times(ms, cnt)
    lib.prs_stage_timing(0)
    clocks = sampler.stop()
    per_bytes, b_alg, passes = algorithmic_bytes(p, sort_interval <= o.timestep)
    peak, peak_src = measured_peak()
    stages = {}
    for i, name in enumerate(STAGES):
        if cnt[i]:
            per_step_us = 1e3 * ms[i] / min(args.steps, 10)
            gbs = per_bytes[name] * n_own / (per_step_us * 1e-6) / 1e9 if per_bytes[name] else None
            stages[name] = {"us_per_step": per_step_us, "alg_bytes_per_robot": per_bytes[name], "achieved_GBps": gbs,
                            "frac_of_hbm_peak": (gbs / peak) if gbs else None}

    # every rank's stage times: the ranks run in lock step through the halo waits, so imbalance shows up as `exchange`
    # time (waiting for a slower neighbour) on the faster ranks
    all_stages = [None] * world
    dist.all_gather_object(all_stages, {k: round(v["us_per_step"], 1) for k, v in stages.items()})
    stages_per_rank = {k: [st_.get(k) for st_ in all_stages] for k in stages}

    # ---- e2e: every rank's pos/vel/rad come from pinned host memory and go back to it every step ----
    e2e_steps = max(3, min(args.steps, 20))
    e2e_s = sim.time_host_steps(o.timestep, sort_interval, e2e_steps)
    finite = torch.tensor([int(torch.isfinite(sim.s.pos[:n_own]).all())], device=dev)
    dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    exchange_used = sim.exchange
    sim.close()
    del sim
    torch.cuda.empty_cache()

    # ---- the same swarm at the same timed state on ONE GPU (rank 0; the other ranks wait) ----
    one = None
    if rank == 0 and not args.no_extras:
        lib.prs_set_stream(C.c_void_p(stream.cuda_stream))
        one = one_gpu_swarm_block(torch, prs, flush, log2n, min(args.steps, 25), evolve_to=evolve_to)
    dist.barrier()
    if rank == 0:
        value = n_total * args.steps / (total_ms_max * 1e-3)
        per_rank = [[int(v) for v in x.tolist()] for x in owns]
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_text(geom, n_total), "decomposition": f"{world} slabs of grid rows",
                       "sort_interval": "timestep (sort every step)",
                       "l2": "flushed between timed steps (256 MiB write)", "halo_rows": multigpu.HALO_ROWS,
                       "exchange": "peer-to-peer stores into the neighbour's mailbox (CUDA IPC over NVLink)" if exchange_used == "p2p"
                                   else "NCCL send/recv of fixed-size buffers"},
            "timed_state_step": evolve_to,
            "scaling_note": "strong scaling of the fixed 2^%d-robot swarm; the driver's --gpus 1 line is S1 (2^20 robots), so the "
                            "same-workload 1-GPU point is one_gpu_same_workload below (measured in this job on rank 0)" % log2n,
            "one_gpu_same_workload": one,
            "speedup_vs_one_gpu": (one["ms_per_step"] / (total_ms_max / args.steps)) if one else None,
            "parity_check": parity,
            "e2e": {"value": n_total * e2e_steps / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 20 * n_total,
                    "d2h_bytes_per_step": 20 * n_total, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "what": "per rank: pos/vel/rad of the owned robots from pinned host -> device, SlabSim.step, device -> pinned host",
                    "cpus_bound_rank0": bound_cpus},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": ({"bound": "hbm", "kernel": "collide (rank 0)", "achieved": stages["collide"]["achieved_GBps"], "peak": peak,
                          "unit": "GB/s", "frac": stages["collide"]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
                          "note": "collide is FP32/MUFU-issue-bound, not HBM-bound; see roofline_step"} if "collide" in stages else None),
            "stages_rank0": stages, "stages_us_per_rank": stages_per_rank, "cpu_baseline": None,
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



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "ref-cuda"])
    ap.add_argument("--robots-log2", type=int, default=None)
    ap.add_argument("--sort-interval", type=float, default=None)
    ap.add_argument("--collide-tile", type=int, default=None, help="1: collide stages neighbour windows in shared memory by TMA; 0: L1/L2 (default: library default)")
    ap.add_argument("--pdl", type=int, default=None, help="1: programmatic dependent launch between the kernels of the fused step (default: library default)")
    ap.add_argument("--try-fused-gather", action="store_true", help="secondary: also time K1+gather as one kernel at this size")
    ap.add_argument("--evolve-to", type=int, default=EVOLVE_TO, help="the timed region starts at this step of the swarm (untimed pre-evolution)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary blocks (pitch 0.155, S2 on one GPU)")
    ap.add_argument("--no-s2", action="store_true", help="skip the 2^26-robot block of the N = 1 line")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the slab-vs-single-GPU bit-equality check before timing")
    ap.add_argument("--sort-only", action="store_true", help="time prs_sort_pairs against cub::DeviceRadixSort on the lattices' cell keys")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-kernels-on-R1 comparison block")
    ap.add_argument("--scramble", action="store_true", help="N = 1: permute the robots so that index order is unrelated to position")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: neighbour exchange by peer-to-peer stores into mapped mailboxes (default) or NCCL send/recv")
    ap.add_argument("--workload", default="s1", choices=["s1", "r1"],
                    help="s1: 2^robots_log2 hex swarm in its own world; r1: 640x640 swarm in the reference's +-64 world")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            run_reference_cpu(args)
        return
    if args.sort_only:
        return run_sort_only(args)
    run_gpu(args)


if __name__ == "__main__":
    main()
