"""SimParams for the CPU oracle WITHOUT the product library.  TEST INFRASTRUCTURE.

`bench.py --impl reference` (the oracle timed on the host cores) must not map libparticlebot_b200.so, so the
defaults of the reference's front end (main.cpp:833-911), its .cfg grammar (main.cpp:594-816, 913-928: a name line
followed by a value line; name lines shorter than 4 characters or starting with '#' are skipped without consuming
a value) and the derived grid (main.cpp:932-939) are restated here in plain Python for the scalar keys the
synthetic workloads use.  tests/test_bench_contract.py checks this against the product's C++ parser on every shipped
cfg.  Only the ctypes structure definition is imported from the package (no library is loaded by that).
"""
import ctypes as C

import numpy as np

from particlerobotsimulations_b200 import SimParams

F32 = np.float32
_keep = []   # obstacle arrays referenced by the structures handed out


def defaults():
    """main.cpp:833-911 (seed is time(NULL) there; 0 here — every shipped cfg sets it)"""
    p = SimParams()
    for name in ("x1obs", "x2obs", "y1obs", "y2obs", "x_cir_obs", "y_cir_obs", "r_cir_obs"):
        arr = (C.c_float * 10)()
        _keep.append(arr)
        setattr(p, name, C.cast(arr, C.POINTER(C.c_float)))
    p.min_radius, p.max_radius = 0.0775, 0.1175
    p.centroid_int, p.centroid_radius, p.centroid_steps = 10, 0.05, 24000
    p.friction, p.spring, p.damping, p.shear = 0.4, 1000.0, 10.0, 40.0
    p.constraint, p.constrained_contraction, p.constraint_contraction = 0.5, 0, 10.0
    p.attraction = float(F32(3.0) * F32(0.000015884))
    p.boundaryDamping = -1.0
    p.gravity = 9.81 * float(F32(0.566))          # double product, then narrowed (main.cpp:866)
    p.nCells, p.nDead = 501, -1
    p.radFactor, p.massFactor, p.frictionFactor, p.attractionFactor = 2.0, 1.0, 1.0, 0.0
    p.time_to_dead, p.max_time, p.seed = 0, 6400.0, 0
    p.light_x, p.light_y, p.light_shadow = -5.0, 0.0, 0
    p.rise_period = 2
    p.phase_std = float(F32(0.3) * F32(2))
    p.config, p.display_shadow, p.phase_update_interval = 0, 0, 12   # CONFIG_RANDOM
    p.control, p.Nx = 0, 5                                               # LIGHT_WAVE
    p.freq = float(F32(0.5) / F32(25))
    run = dict(timestep=float(F32(0.01)), sort_interval=180.0, dump_interval=60.0)
    derive_grid(p)
    return p, run


def derive_grid(p):
    """main.cpp:932-939"""
    if p.nDead == -1 and p.max_radius * 0.5 * p.radFactor > 2 * p.max_radius:
        cell = p.max_radius * 0.5 * p.radFactor + 4 * p.max_radius
    else:
        cell = p.max_radius * 2
    p.cellSize.x = p.cellSize.y = cell
    set_world(p, 512, 64.0)


def set_world(p, grid_dim, world_half):
    """synthetic worlds of SURVEY.md §8d (prs_params_set_world)"""
    p.gridSize.x = p.gridSize.y = grid_dim
    p.numCells = grid_dim * grid_dim
    p.worldOrigin.x = p.worldOrigin.y = -world_half


_FLOAT = {"min_radius", "max_radius", "centroid_radius", "radFactor", "massFactor", "frictionFactor", "attractionFactor",
          "friction", "spring", "damping", "shear", "constraint", "attraction", "boundaryDamping", "gravity",
          "time_to_dead", "max_time", "light_x", "light_y", "rise_period", "phase_std"}
_INT = {"centroid_int", "centroid_steps", "testing", "constrained_contraction", "nCells", "nDead", "seed", "light_shadow",
        "display_shadow", "phase_update_interval"}
_RUN = {"timestep", "sort_interval", "dump_interval"}


def load_cfg(path):
    """scalar keys only (obstacle lists are not needed by the synthetic workloads: refused loudly)"""
    p, run = defaults()
    lines = open(path).read().split("\n")
    i = 0
    while i < len(lines):
        name = lines[i]
        i += 1
        if len(name) < 4 or name.startswith("#"):
            continue
        if i >= len(lines):
            break
        value = lines[i]
        i += 1
        key = name.strip()
        if key in _FLOAT:
            setattr(p, key, float(F32(value)))
        elif key in _INT:
            setattr(p, key, int(value))
        elif key in _RUN:
            run[key] = float(F32(value))
        elif "obs" in key:
            raise ValueError(f"oracle/params.py does not parse obstacle lists ({key}); use the product parser")
    derive_grid(p)
    return p, run
