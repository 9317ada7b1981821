"""ctypes binding of oracle/_build/libprs_oracle.so (see oracle/prs_oracle.h).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

from particlerobotsimulations_b200 import SimParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libprs_oracle.so")
REFCUDA_PATH = os.path.join(_HERE, "_ref", "libprs_refcuda.so")
REFHOST_PATH = os.path.join(_HERE, "_ref", "libprs_refhost.so")
DROPIN_PATH = os.path.join(_HERE, "_ref", "libprs_dropin.so")   # the reference's host class over the PRODUCT library
# ... and over the product's OpenGL build (-DPRS_WITH_GL: its own VBO interop entry points, CUDA graphics API -> headless buffer objects)
DROPIN_GL_PATH = os.path.join(_HERE, "_ref", "libprs_dropin_gl.so")


def build():
    """Compile the CPU restatement and, when /root/reference is present, the verbatim reference kernels."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


class RngState(C.Structure):
    _fields_ = [("d", C.c_uint), ("v", C.c_uint * 5), ("boxmuller_flag", C.c_int), ("boxmuller_flag_double", C.c_int),
                ("boxmuller_extra", C.c_float), ("pad_", C.c_int), ("boxmuller_extra_double", C.c_double)]


assert C.sizeof(RngState) == 48


class GlibcRand(C.Structure):
    _fields_ = [("r", C.c_int * 34), ("f", C.c_int), ("b", C.c_int)]


_lib = None
_VP, _F, _I, _U = C.c_void_p, C.c_float, C.c_int, C.c_uint
_PP = C.POINTER(SimParams)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        sig = {
            "prso_set_threads": (None, [_I]), "prso_get_max_threads": (_I, []),
            "prso_srand": (None, [C.POINTER(GlibcRand), _U]), "prso_rand": (_I, [C.POINTER(GlibcRand)]),
            "prso_calc_hash": (None, [_PP, _VP, _VP, _VP, _I]),
            "prso_sort_pairs": (None, [_VP, _VP, _I]),
            "prso_reorder_find_cell_start": (None, [_PP] + [_VP] * 10 + [_I, _U]),
            "prso_collide": (None, [_PP] + [_VP] * 9 + [_I, _F]),
            "prso_integrate": (None, [_PP, _VP, _VP, _VP, _F, _I, _F]),
            "prso_update_rad": (None, [_PP, _VP, _VP, _VP, _VP, _F, _F, _VP, _I]),
            "prso_update_phase": (None, [_PP, _VP, _VP, _F, _F, _I]),
            "prso_min_light_distance": (_F, [_PP, _VP, _I]),
            "prso_curand_setup": (None, [_VP, _U, _I]), "prso_add_normal_noise": (None, [_VP, _VP, _F, _I]),
            "prso_create": (_VP, [_PP, _F]), "prso_destroy": (None, [_VP]), "prso_srand_sim": (None, [_VP, _U]),
            "prso_reset": (None, [_VP]), "prso_update": (None, [_VP, _F, _F]), "prso_time": (_F, [_VP]),
            "prso_array": (_VP, [_VP, _I]), "prso_gate": (_I, [_F, _F, _F]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def ptr(a):
    return a.ctypes.data


_ARR = {"pos": (0, np.float32, 2), "vel": (1, np.float32, 2), "rad": (2, np.float32, 1), "phase": (3, np.float32, 1),
        "absForce_a": (4, np.float32, 1), "absForce_r": (5, np.float32, 1), "dead": (6, np.int32, 1),
        "hash": (7, np.uint32, 1), "index": (8, np.uint32, 1), "cellStart": (9, np.uint32, 0),
        "cellEnd": (10, np.uint32, 0), "sortedPos": (11, np.float32, 2), "sortedVel": (12, np.float32, 2),
        "sortedRad": (13, np.float32, 1)}


class OracleSim:
    """Particlebot::{ctor, reset, update} on the CPU (prs_oracle.cpp)."""

    def __init__(self, params, world_half=64.0):
        self.L = lib()
        self.params = params
        self.n = int(params.nCells)
        self.h = self.L.prso_create(C.byref(params), world_half)

    def close(self):
        if self.h:
            self.L.prso_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def srand(self, seed):
        self.L.prso_srand_sim(self.h, seed)

    def reset(self):
        self.L.prso_reset(self.h)

    def update(self, dt, sort_interval):
        self.L.prso_update(self.h, dt, sort_interval)

    @property
    def time(self):
        return float(self.L.prso_time(self.h))

    def view(self, name):
        """numpy VIEW of an oracle buffer (writes go through)."""
        which, dt, w = _ARR[name]
        count = int(self.params.numCells) if w == 0 else self.n * w
        p = self.L.prso_array(self.h, which)
        buf = (C.c_char * (count * np.dtype(dt).itemsize)).from_address(p)
        a = np.frombuffer(buf, dtype=dt)
        return a.reshape(self.n, 2) if w == 2 else a

    def get(self, name):
        return self.view(name).copy()


_refhost = {}


def refhost(path=None):
    """oracle/_ref/libprs_refhost.so: the reference's own host class (particlebot.cpp) and kernels compiled verbatim,
    OpenGL buffer objects replaced by device allocations (oracle/gl_stub).  path = DROPIN_PATH: the same class linked
    against libparticlebot_b200.so instead of the reference's kernels.  Needs a GPU."""
    path = path or REFHOST_PATH
    if path not in _refhost:
        L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        sig = {"prsref_identity": (C.c_char_p, []), "prsref_create": (_VP, [_PP, _U]), "prsref_destroy": (None, [_VP]),
               "prsref_reset": (None, [_VP]), "prsref_update": (None, [_VP, _F, _F]), "prsref_time": (_F, [_VP]),
               "prsref_get": (_I, [_VP, _I, _VP, C.c_size_t]), "prsref_set": (None, [_VP, _I, _VP, _I, _I])}
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _refhost[path] = L
    return _refhost[path]


class RefHostSim:
    """`class Particlebot` of the REFERENCE (particlebot.cpp), driven like main.cpp does: srand(seed), construct, reset(),
    update(dt, sort_interval) per step.  It uses the process-wide glibc rand(): one at a time."""

    _GET = {"pos": (0, np.float32, 2), "vel": (1, np.float32, 2), "rad": (2, np.float32, 1), "phase": (3, np.float32, 1),
            "dead": (5, np.int32, 1), "absForce_a": (100, np.float32, 1), "absForce_r": (101, np.float32, 1),
            "hash": (102, np.uint32, 1), "index": (103, np.uint32, 1), "cellStart": (104, np.uint32, 0), "cellEnd": (105, np.uint32, 0)}

    def __init__(self, params, seed, library=None):
        self.L = refhost(library)
        self.params = params
        self.n = int(params.nCells)
        self.h = self.L.prsref_create(C.byref(params), seed)

    def close(self):
        if self.h:
            self.L.prsref_destroy(self.h)
            self.h = None

    def reset(self):
        self.L.prsref_reset(self.h)

    def update(self, dt, sort_interval):
        self.L.prsref_update(self.h, dt, sort_interval)

    @property
    def time(self):
        return float(self.L.prsref_time(self.h))

    def get(self, name):
        which, dt, w = self._GET[name]
        shape = (int(self.params.numCells),) if w == 0 else ((self.n, 2) if w == 2 else (self.n,))
        out = np.empty(shape, dt)
        rc = self.L.prsref_get(self.h, which, out.ctypes.data, out.nbytes)
        if rc != 0:
            raise RuntimeError(f"prsref_get({name}) failed: {rc}")
        return out
