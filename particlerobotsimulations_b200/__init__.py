"""particlerobotsimulations_b200 — ctypes view of libparticlebot_b200.so.

The product is the C++/CUDA library (csrc/, include/): hand-written sm_100a kernels behind the
reference's own `extern "C"` entry points plus the headless `Particlebot` class.  This module
only loads it for tests, bench.py and Python users; there is NO CPU fallback — if the library is
missing or no CUDA device is present, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PRS_LIB: another build of the SAME library (A/B measurements of compile-time variants, scripts/build_variant.sh)
LIB_PATH = os.environ.get("PRS_LIB") or os.path.join(_HERE, "libparticlebot_b200.so")

# array selectors of prs_sim_get/set/device_ptr (ParticlebotArray + additions, prs_cabi.h)
POSITION, VELOCITY, RADII, PHASE, FREQUENCY, DEAD = range(6)
ABSFORCE_A, ABSFORCE_R, HASH, INDEX, CELLSTART, CELLEND, SORTEDPOS, SORTEDVEL, SORTEDRAD, RNGSTATE = range(100, 110)
BACKEND_FUSED, BACKEND_PERCALL, BACKEND_EXTERNAL = 0, 1, 2


class _U2(C.Structure):
    _fields_ = [("x", C.c_uint), ("y", C.c_uint)]


class _F2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class SimParams(C.Structure):
    """Mirror of include/prs_simparams.h (reference particlebot_kernel.cuh:58-120), 256 bytes."""

    _fields_ = [
        ("gridSize", _U2), ("numCells", C.c_uint), ("_pad0", C.c_uint),
        ("worldOrigin", _F2), ("cellSize", _F2),
        ("nCells", C.c_uint), ("nDead", C.c_int), ("maxParticlebotsPerCell", C.c_uint),
        ("gravity", C.c_float), ("spring", C.c_float), ("damping", C.c_float), ("shear", C.c_float),
        ("attraction", C.c_float), ("boundaryDamping", C.c_float), ("friction", C.c_float),
        ("massFactor", C.c_float), ("frictionFactor", C.c_float), ("radFactor", C.c_float),
        ("attractionFactor", C.c_float), ("constraint", C.c_float), ("constraint_contraction", C.c_float),
        ("centroid_steps", C.c_int), ("centroid_int", C.c_float), ("centroid_radius", C.c_float),
        ("light_x", C.c_float), ("light_y", C.c_float), ("phase_update_interval", C.c_float),
        ("control", C.c_int), ("config", C.c_int),
        ("min_radius", C.c_float), ("max_radius", C.c_float), ("rise_period", C.c_float), ("freq", C.c_float),
        ("nobstacles", C.c_int),
        ("x1obs", C.POINTER(C.c_float)), ("x2obs", C.POINTER(C.c_float)),
        ("y1obs", C.POINTER(C.c_float)), ("y2obs", C.POINTER(C.c_float)),
        ("n_cir_obstacles", C.c_int),
        ("x_cir_obs", C.POINTER(C.c_float)), ("y_cir_obs", C.POINTER(C.c_float)), ("r_cir_obs", C.POINTER(C.c_float)),
        ("Nx", C.c_int), ("phase_std", C.c_float), ("seed", C.c_uint),
        ("light_shadow", C.c_uint), ("testing", C.c_uint), ("constrained_contraction", C.c_uint),
        ("display_shadow", C.c_uint), ("time_to_dead", C.c_float), ("max_time", C.c_float),
    ]


assert C.sizeof(SimParams) == 256, C.sizeof(SimParams)


class RunOptions(C.Structure):
    _fields_ = [
        ("timestep", C.c_float), ("sort_interval", C.c_float), ("dump_interval", C.c_float),
        ("camera_x", C.c_float), ("camera_y", C.c_float), ("light_radius", C.c_float),
        ("display_interval", C.c_int), ("video_interval", C.c_int),
        ("csv_filename", C.c_char * 300), ("video_filename", C.c_char * 300),
        ("init_hexblock", C.c_int), ("hexblock_nx", C.c_uint), ("hexblock_ny", C.c_uint), ("hexblock_seed", C.c_uint),
        ("hexblock_pitch", C.c_float), ("hexblock_jitter", C.c_float), ("world_half", C.c_float), ("grid_dim", C.c_uint),
    ]


class View(C.Structure):
    """Mirror of prs_view (include/prs_cabi.h): the straight-down camera of a headless frame."""

    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("center_x", C.c_float), ("center_y", C.c_float),
                ("world_per_pixel", C.c_float), ("light_radius", C.c_float)]


class StepBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "pos", "vel", "rad", "phase", "absForce_a", "absForce_r", "dead", "hash", "index", "cellStart", "cellEnd",
        "sortedPos", "sortedVel", "sortedRad")] + [("nCells", C.c_uint), ("numCells", C.c_uint), ("sortedPR", C.c_void_p)]


class Slab(C.Structure):
    """Mirror of prs_slab (include/prs_cabi.h): one rank's slab of the swarm, device pointers."""

    _fields_ = [(n, C.c_void_p) for n in (
        "pos", "vel", "rad", "phase", "absForce_a", "absForce_r", "dead", "gid", "rng", "hash", "scratch",
        "sortedPR", "sortedVel", "hash_cat", "index_sorted", "cellStart", "cellEnd", "counts", "lists")] + [
        (n, C.c_uint) for n in ("cap", "halo_cap", "mig_cap", "row_lo", "row_hi", "halo_rows")] + [
        ("has_dn", C.c_int), ("has_up", C.c_int), ("wrap", C.c_int)]


ALLREDUCE_MIN_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)


class SlabCtx(C.Structure):
    """Mirror of prs_slab_ctx (include/prs_cabi.h): a slab rank's whole-step context for prs_slab_step."""

    _fields_ = [("slab", Slab), ("mailbox", C.c_void_p), ("peer_dn", C.c_void_p), ("peer_up", C.c_void_p),
                ("mw", C.c_uint), ("hw", C.c_uint), ("scratch_mig", C.c_void_p * 2), ("scratch_halo", C.c_void_p * 2),
                ("d_min_d", C.c_void_p), ("allreduce_min", ALLREDUCE_MIN_FN), ("user", C.c_void_p), ("overlap_exchange", C.c_int),
                ("time", C.c_float), ("sorted_once", C.c_int), ("seq_halo", C.c_uint), ("seq_mig", C.c_uint),
                ("split_fallbacks", C.c_uint), ("h_err", C.c_void_p), ("fused_exchange", C.c_int)]


# words of Slab.counts (PRS_SC_*) and error bits (PRS_SLAB_ERR_*)
SC_N, SC_NLO, SC_NHI, SC_KDN, SC_KUP, SC_MIGDN, SC_MIGUP, SC_LEAVERS, SC_HOLES, SC_KEEPERS, SC_ERR, SC_STAT_MIG, SC_STAT_HALO = range(13)
SLAB_ERRORS = {1: "more migrants than mig_cap in one step", 2: "a halo longer than halo_cap", 4: "more robots than the slab capacity",
               8: "a robot crossed more than one slab between two sorts", 16: "a robot left the grid rows of the outermost slab",
               32: "peer-to-peer exchange timed out waiting for a neighbour",
               64: "an owned robot drifted beyond the halo rows between two sorts (raise halo_rows or sort more often)"}
SLAB_MIG_WORDS, SLAB_HALO_WORDS = 23, 7

_lib = None

# name -> (restype, argtypes); the reference's entry points first (include/prs_cabi.h part 1)
_VP, _F, _I, _U = C.c_void_p, C.c_float, C.c_int, C.c_uint
SIGNATURES = {
    "cudaInit": (None, [_I, _VP]), "cudaGLInit": (None, [_I, _VP]),
    "allocateArray": (None, [C.POINTER(_VP), C.c_size_t]), "freeArray": (None, [_VP]), "threadSync": (None, []),
    "copyArrayToDevice": (None, [_VP, _VP, _I, _I]), "copyArrayFromDevice": (None, [_VP, _VP, _VP, _I]),
    "registerGLBufferObject": (None, [_U, _VP]), "unregisterGLBufferObject": (None, [_VP]),
    "mapGLBufferObject": (_VP, [_VP]), "unmapGLBufferObject": (None, [_VP]),
    "setParameters": (None, [C.POINTER(SimParams)]), "iDivUp": (_U, [_U, _U]),
    "integrateSystem": (None, [_VP, _VP, _VP, _F, _U, _F]),
    "calcHash": (None, [_VP, _VP, _VP, _I]),
    "sortParticlebots": (None, [_VP, _VP, _U]),
    "reorderDataAndFindCellStart": (None, [_VP] * 10 + [_U, _U]),
    "collide": (None, [_VP] * 9 + [_U, _U, _F]),
    "updateRad_light_wave": (None, [_VP, _VP, _VP, _VP, _VP, _F, _F, _VP, _I]),
    "updatePhase": (None, [_VP, _VP, _F, _F, _F, _I]),
    "curand_setup": (None, [_VP, _I]), "add_normal_noise": (None, [_VP, _VP, _F, _I]),
    "calcCOG": (None, [_VP, _VP, _VP, _I, _F, _I, _F]),
    "updateCol": (None, [_VP, _VP, _I, _VP, _VP, _VP]),
    # part 2
    "prs_version": (C.c_char_p, []), "prs_set_stream": (None, [_VP]), "prs_get_stream": (_VP, []),
    "prs_set_world_half_extent": (None, [_F]), "prs_get_world_half_extent": (_F, []),
    "prs_set_collide_warp_max": (None, [_U]),
    "prs_set_collide_tile": (None, [_I]), "prs_get_collide_tile": (_I, []),
    "prs_set_patch_rows": (None, [_U]), "prs_patch_stats": (None, [_I, _VP]),
    "prs_set_pdl": (None, [_I]), "prs_get_pdl": (_I, []), "prs_set_k1_x2": (None, [_I]), "prs_set_collide_dense": (None, [_I]), "prs_set_slab_scan_range": (None, [_I]),
    "prs_set_fuse_gather_max": (None, [_U]),
    "prs_launch_count": (C.c_ulonglong, [_I]),
    "prs_stage_timing": (None, [_I]), "prs_stage_times": (None, [_VP, _VP]),
    "prs_min_light_distance": (None, [_VP, _I, _VP]), "prs_update_phase_dev": (None, [_VP, _VP, _F, _VP, _I]),
    "prs_centroid": (None, [_VP, _I, _VP, _VP]),
    "prs_sort_pairs": (None, [_VP, _VP, _VP, _VP, _U, _I]),
    "prs_sort_set_timeline": (None, [_VP]), "prs_sort_tile_size": (_U, []), "prs_sort_plan": (_I, [_I, _U, _VP]), "prs_sort_set_threads": (None, [_I]),
    "prs_slab_mig_words": (C.c_size_t, [_U]), "prs_slab_halo_words": (C.c_size_t, [_U]),
    "prs_slab_rng_setup": (None, [_VP, _U]), "prs_slab_k1": (None, [_VP, _F, _F, _I]),
    "prs_slab_migrate_pack": (None, [_VP, _VP, _VP]), "prs_slab_migrate_unpack": (None, [_VP, _VP, _VP]),
    "prs_slab_sort": (None, [_VP]), "prs_slab_gather": (None, [_VP]),
    "prs_slab_halo_pack": (None, [_VP, _VP, _VP]), "prs_slab_halo_unpack": (None, [_VP, _VP, _VP]),
    "prs_slab_cell_table": (None, [_VP]), "prs_slab_collide": (None, [_VP, _F]),
    "prs_slab_min_light_distance": (None, [_VP, _VP]), "prs_slab_update_phase": (None, [_VP, _F, _VP]),
    "prs_slab_add_noise": (None, [_VP, _F]),
    "prs_ipc_handle_size": (C.c_size_t, []), "prs_slab_mailbox_alloc": (_VP, [C.c_size_t]), "prs_slab_mailbox_free": (None, [_VP]),
    "prs_ipc_export": (None, [_VP, _VP]), "prs_ipc_open": (_VP, [_VP]), "prs_ipc_close": (None, [_VP]),
    "prs_slab_signal": (None, [_VP, _VP, _U]), "prs_slab_wait": (None, [_VP, _VP, _VP, _U]),
    "prs_slab_mailbox_words": (C.c_size_t, [_U, _U]), "prs_slab_step": (_U, [_VP, _F, _F]), "prs_slab_ctx_release": (None, [_VP]),
    "prs_init_hex_block": (None, [_VP, _VP, _VP, _VP, _VP, _U, _U, _U, _F, _F, _U, _F]),
    "prs_unpack_sorted": (None, [_VP, _VP, _VP, _U]), "prs_selftest_div": (C.c_ulonglong, [_VP, _VP, _U]),
    "prs_fused_step": (None, [C.POINTER(StepBuffers), _F, _F, _I]),
    "prs_bin_invalidate": (None, []), "prs_bin_set_mode": (None, [_I]), "prs_bin_active": (_I, []),
    "prs_params_defaults": (None, [C.POINTER(SimParams), C.POINTER(RunOptions)]),
    "prs_params_load_cfg": (_I, [C.c_char_p, C.POINTER(SimParams), C.POINTER(RunOptions)]),
    "prs_params_derive_grid": (None, [C.POINTER(SimParams)]),
    "prs_params_set_world": (None, [C.POINTER(SimParams), _U, _F]),
    "prs_sim_create": (_VP, [C.POINTER(SimParams), _F, _I, C.c_char_p]), "prs_sim_destroy": (None, [_VP]),
    "prs_sim_srand": (None, [_VP, _U]), "prs_sim_reset": (None, [_VP]),
    "prs_sim_init_hex": (None, [_VP, _U, _U, _F, _F, _U]),
    "prs_sim_update": (_I, [_VP, _F, _F]), "prs_sim_time": (_F, [_VP]), "prs_sim_sync": (None, [_VP]),
    "prs_sim_device_ptr": (_VP, [_VP, _I]), "prs_sim_get": (None, [_VP, _I, _VP, C.c_size_t]),
    "prs_sim_set": (None, [_VP, _I, _VP, C.c_size_t, C.c_size_t]),
    "prs_sim_dump": (None, [_VP, _VP, _F, _U]), "prs_sim_load": (None, [_VP, _VP]),
    "prs_sim_update_host": (_I, [_VP] * 7 + [_F, _F]),
    "prs_h2d_async": (None, [_VP, _VP, C.c_size_t]), "prs_arm_k1_event": (None, [_I]),
    "prs_d2h_async": (None, [_VP, _VP, C.c_size_t, _I]), "prs_host_step_sync": (None, []),
    "prs_host_step_plan": (None, [_VP] * 5), "prs_set_plan_chunks": (None, [_U]),
    "prs_sim_checkpoint_save": (_I, [_VP, C.c_char_p]), "prs_sim_checkpoint_load": (_I, [_VP, C.c_char_p]),
    # headless frames and video
    "prs_view_from_camera": (None, [C.POINTER(View), _U, _U, _F, _F]),
    "prs_render_frame": (None, [_VP, _VP, C.POINTER(View), _VP, _VP, _VP, _U]),
    "prs_sim_render_frame": (_VP, [_VP, C.POINTER(View)]),
    "prs_video_open": (_VP, [C.c_char_p, _U, _U, C.c_double]), "prs_video_write": (_I, [_VP, _VP]), "prs_video_close": (_I, [_VP]),
}


def bind_signatures(lib, names=None):
    for name, (res, args) in SIGNATURES.items():
        if names is not None and name not in names:
            continue
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def lib():
    """The product library.  Raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python particlerobotsimulations_b200/build.py` "
                "(nvcc, sm_100a).  There is no CPU or PyTorch fallback.")
        _lib = bind_signatures(C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW))
    return _lib


def view_from_camera(width, height, camera_y, light_radius):
    v = View()
    lib().prs_view_from_camera(C.byref(v), width, height, camera_y, light_radius)
    return v


class VideoWriter:
    """prs_video_* (csrc/prs_video.cpp): uncompressed AVI of top-down B, G, R frames; host code only"""

    def __init__(self, path, width, height, fps=20.0):
        self._lib = lib()
        self.shape = (height, width, 3)
        self._h = self._lib.prs_video_open(os.fsencode(path), width, height, fps)
        if not self._h:
            raise OSError(f"cannot open {path}")

    def write(self, frame):
        f = np.ascontiguousarray(frame, np.uint8)
        assert f.shape == self.shape, (f.shape, self.shape)
        return self._lib.prs_video_write(self._h, f.ctypes.data)

    def close(self):
        n, self._h = self._lib.prs_video_close(self._h), None
        return n


def default_params():
    p, o = SimParams(), RunOptions()
    lib().prs_params_defaults(C.byref(p), C.byref(o))
    return p, o


def load_cfg(path):
    """Defaults (main.cpp:833-911) + .cfg file (main.cpp:594-816, 913-928) + derived grid (:932-939)."""
    p, o = default_params()
    rc = lib().prs_params_load_cfg(os.fsencode(path), C.byref(p), C.byref(o))
    if rc != 0:
        raise FileNotFoundError(path)
    return p, o


_DTYPES = {POSITION: (np.float32, 2), VELOCITY: (np.float32, 2), RADII: (np.float32, 1), PHASE: (np.float32, 1),
           FREQUENCY: (np.float32, 1), DEAD: (np.int32, 1), ABSFORCE_A: (np.float32, 1), ABSFORCE_R: (np.float32, 1),
           HASH: (np.uint32, 1), INDEX: (np.uint32, 1), SORTEDPOS: (np.float32, 2), SORTEDVEL: (np.float32, 2),
           SORTEDRAD: (np.float32, 1), RNGSTATE: (np.uint32, 12)}   # 48-byte curandStateXORWOW per robot


class Simulation:
    """`class Particlebot` (include/prs_particlebot.hpp) through its C wrappers."""

    def __init__(self, params, world_half=64.0, backend=BACKEND_FUSED, external_library=None):
        self._lib = lib()
        self.params = params
        self.n = int(params.nCells)
        ext = os.fsencode(external_library) if external_library else None
        self._h = self._lib.prs_sim_create(C.byref(params), world_half, backend, ext)

    def close(self):
        if self._h:
            self._lib.prs_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def srand(self, seed):
        self._lib.prs_sim_srand(self._h, seed)

    def reset(self):
        self._lib.prs_sim_reset(self._h)

    def init_hex(self, nx, ny, pitch, jitter, seed):
        self._lib.prs_sim_init_hex(self._h, nx, ny, pitch, jitter, seed)

    def update(self, dt, sort_interval):
        return bool(self._lib.prs_sim_update(self._h, dt, sort_interval))

    def sync(self):
        self._lib.prs_sim_sync(self._h)

    def update_host(self, pos, vel, rad, dt, sort_interval):
        """one step with the state in host arrays (float32, C-contiguous, ideally pinned): updated in place"""
        return bool(self._lib.prs_sim_update_host(self._h, pos.ctypes.data, vel.ctypes.data, rad.ctypes.data,
                                                  pos.ctypes.data, vel.ctypes.data, rad.ctypes.data, dt, sort_interval))

    def save_checkpoint(self, path):
        """full binary checkpoint (Particlebot::saveCheckpoint); raises OSError on failure"""
        if self._lib.prs_sim_checkpoint_save(self._h, os.fsencode(path)) != 0:
            raise OSError(f"cannot write checkpoint {path}")

    def load_checkpoint(self, path):
        """restore a checkpoint of a swarm of the same shape; the run continues bit for bit"""
        if self._lib.prs_sim_checkpoint_load(self._h, os.fsencode(path)) != 0:
            raise OSError(f"cannot restore checkpoint {path}")

    def render_frame(self, view):
        """one headless frame (Particlebot::renderFrame): uint8 array (height, width, 3), top-down rows, B, G, R"""
        ptr = self._lib.prs_sim_render_frame(self._h, C.byref(view))
        h, w = int(view.height), int(view.width)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_ubyte)), shape=(h, w, 3)).copy()

    @property
    def time(self):
        return float(self._lib.prs_sim_time(self._h))

    def device_ptr(self, which):
        return self._lib.prs_sim_device_ptr(self._h, which)

    def get(self, which):
        if which in (CELLSTART, CELLEND):
            out = np.empty(int(self.params.numCells), np.uint32)
        else:
            dt, w = _DTYPES[which]
            out = np.empty((self.n, w) if w > 1 else (self.n,), dt)
        self._lib.prs_sim_get(self._h, which, out.ctypes.data, out.nbytes)
        return out

    def set(self, which, arr, offset_items=0):
        dt, w = _DTYPES[which]
        a = np.ascontiguousarray(arr, dt)
        self._lib.prs_sim_set(self._h, which, a.ctypes.data, offset_items * a.itemsize * w, a.nbytes)
