"""Helpers shared by the tests: cfg loading, device buffers through the C-ABI, oracle access."""
import ctypes as C
import os

import numpy as np

import particlerobotsimulations_b200 as prs
from oracle import binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")
GOLDEN = os.path.join(ROOT, "tests", "golden")
CFGS = ["example", "example_dead_cells", "example_obstacle", "example_gap", "example_object_transport"]


def cfg(name):
    return prs.load_cfg(os.path.join(EXAMPLES, name + ".cfg"))


def refcuda_available():
    return os.path.exists(ob.REFCUDA_PATH)


def refhost_available():
    return os.path.exists(ob.REFHOST_PATH)


_ref = None


def refcuda():
    """The reference's own kernels + extern "C" wrappers compiled verbatim (oracle/_ref)."""
    global _ref
    if _ref is None:
        L = C.CDLL(ob.REFCUDA_PATH, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        names = ["allocateArray", "freeArray", "threadSync", "copyArrayToDevice", "copyArrayFromDevice",
                 "setParameters", "integrateSystem", "calcHash", "sortParticlebots", "reorderDataAndFindCellStart",
                 "collide", "updateRad_light_wave", "updatePhase", "curand_setup", "add_normal_noise", "calcCOG", "updateCol"]
        prs.bind_signatures(L, names)
        _ref = L
    return _ref


class Dev:
    """A device buffer owned through the C-ABI (allocateArray / copyArrayToDevice / copyArrayFromDevice)."""

    def __init__(self, arr_or_bytes, dtype=None, lib=None):
        self.lib = lib or prs.lib()
        if isinstance(arr_or_bytes, (int, np.integer)):
            self.nbytes = int(arr_or_bytes)
            self.dtype, self.shape = dtype or np.uint8, None
            host = None
        else:
            host = np.ascontiguousarray(arr_or_bytes)
            self.nbytes, self.dtype, self.shape = host.nbytes, host.dtype, host.shape
        p = C.c_void_p()
        self.lib.allocateArray(C.byref(p), max(self.nbytes, 16))
        self.ptr = p.value
        if host is not None and self.nbytes:
            self.lib.copyArrayToDevice(self.ptr, host.ctypes.data, 0, self.nbytes)

    def get(self, dtype=None, shape=None):
        dtype = dtype or self.dtype
        out = np.empty(self.nbytes // np.dtype(dtype).itemsize, dtype)
        if self.nbytes:
            self.lib.copyArrayFromDevice(out.ctypes.data, self.ptr, None, self.nbytes)
        shape = shape or self.shape
        return out.reshape(shape) if shape else out

    def set(self, arr):
        a = np.ascontiguousarray(arr)
        assert a.nbytes == self.nbytes
        self.lib.copyArrayToDevice(self.ptr, a.ctypes.data, 0, a.nbytes)

    def free(self):
        if self.ptr:
            self.lib.freeArray(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def oracle_state_after(params, opt, steps, sort_interval=None, world_half=64.0):
    s = ob.OracleSim(params, world_half)
    s.srand(params.seed)
    s.reset()
    si = opt.sort_interval if sort_interval is None else sort_interval
    for _ in range(steps):
        s.update(opt.timestep, si)
    return s


def rel_err(a, b, floor):
    """max |a-b| / max(|b|, floor): elementwise relative error with an absolute floor"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
