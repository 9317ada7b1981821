"""GPU tests (-m gpu) against the REFERENCE'S OWN HOST CLASS: oracle/_ref/libprs_refhost.so is particlebot.cpp and
particlebot_cuda.cu compiled verbatim (buffer objects replaced by device allocations, oracle/gl_stub), driven the way
main.cpp:929-952 + 360-361 drives it — srand(seed), Particlebot(params), reset(), update(timestep, sort_interval) per
step.  What is pinned here is the HOST logic of this repository (csrc/prs_particlebot.cpp and, through the goldens,
oracle/prs_oracle.cpp): CONFIG_RANDOM placement (particlebot.cpp:612-748), radii / dead flags / phases of reset()
(:775-800), the update order and its fp32 gates (:170-300), the dead-cell draw on the glibc stream (:178-194) and the
host min-distance loop (:214-228).  The kernels are the same on both sides in the EXTERNAL backend (identical bits
expected); the fused backend replaces them with this library's own."""
import os

import numpy as np
import pytest

import particlerobotsimulations_b200 as prs
from oracle import binding as ob
from tests import util

pytestmark = pytest.mark.gpu

STEPS = (1, 10, 50, 100)


def _need_refhost():
    if not util.refhost_available():
        pytest.skip("oracle/_ref/libprs_refhost.so not built (needs /root/reference at build time)")


def _reference_run(name, sort_every_step, library=None):
    p, o = util.cfg(name)
    ref = ob.RefHostSim(p, p.seed, library)
    ref.reset()
    snaps = {0: {k: ref.get(k) for k in ("pos", "rad", "dead", "phase")}}
    si = o.timestep if sort_every_step else o.sort_interval
    for k in range(1, max(STEPS) + 1):
        ref.update(o.timestep, si)
        if k in STEPS:
            snaps[k] = {key: ref.get(key) for key in ("pos", "vel", "rad", "phase", "dead", "hash", "index", "cellStart", "cellEnd",
                                                      "absForce_r", "absForce_a")}
    assert abs(ref.time - max(STEPS) * o.timestep) < 1e-3
    ref.close()
    return p, o, snaps


def _own_run(p, o, backend, ext, sort_every_step):
    sim = prs.Simulation(p, 64.0, backend, ext)
    sim.srand(p.seed)
    sim.reset()
    snaps = {0: {"pos": sim.get(prs.POSITION), "rad": sim.get(prs.RADII), "dead": sim.get(prs.DEAD), "phase": sim.get(prs.PHASE)}}
    si = o.timestep if sort_every_step else o.sort_interval
    for k in range(1, max(STEPS) + 1):
        sim.update(o.timestep, si)
        if k in STEPS:
            snaps[k] = {"pos": sim.get(prs.POSITION), "vel": sim.get(prs.VELOCITY), "rad": sim.get(prs.RADII),
                        "phase": sim.get(prs.PHASE), "dead": sim.get(prs.DEAD), "hash": sim.get(prs.HASH),
                        "index": sim.get(prs.INDEX), "cellStart": sim.get(prs.CELLSTART), "cellEnd": sim.get(prs.CELLEND),
                        "absForce_r": sim.get(prs.ABSFORCE_R), "absForce_a": sim.get(prs.ABSFORCE_A)}
    sim.close()
    return snaps


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", util.CFGS)
@pytest.mark.parametrize("sort_every_step", [False, True])
def test_host_logic_equals_the_reference_class(name, sort_every_step, collide_kernel_default):
    """this repository's Particlebot over the reference's kernels (EXTERNAL backend) == the reference's Particlebot over
    the same kernels, bit for bit: placement, dead draw, phases with noise, 100 steps in both sort cadences"""
    _need_refhost()
    if not util.refcuda_available():
        pytest.skip("oracle/_ref/libprs_refcuda.so not built")
    p, o, ref = _reference_run(name, sort_every_step)
    own = _own_run(p, o, prs.BACKEND_EXTERNAL, ob.REFCUDA_PATH, sort_every_step)
    for key in ("pos", "rad", "dead", "phase"):
        assert np.array_equal(_bits(own[0][key]), _bits(ref[0][key])), ("after reset()", key)
    for k in STEPS:
        occ = np.nonzero(ref[k]["cellStart"] != 0xFFFFFFFF)[0]
        assert np.array_equal(np.nonzero(own[k]["cellStart"] != 0xFFFFFFFF)[0], occ), k
        for key in ("dead", "hash", "index", "pos", "vel", "rad", "phase", "absForce_r", "absForce_a"):
            assert np.array_equal(_bits(own[k][key]), _bits(ref[k][key])), (k, key)
        assert np.array_equal(own[k]["cellStart"][occ], ref[k]["cellStart"][occ]) and np.array_equal(own[k]["cellEnd"][occ], ref[k]["cellEnd"][occ]), k


@pytest.mark.parametrize("name", util.CFGS)
@pytest.mark.parametrize("sort_every_step", [False, True])
def test_fused_path_vs_the_reference_class(name, sort_every_step, collide_kernel_default):
    """the product path (fused step, own kernels, device-side min distance) against the reference's class: integer state
    bit-exact, floats at the north_star bar (1e-5 relative over 100 steps)"""
    _need_refhost()
    p, o, ref = _reference_run(name, sort_every_step)
    own = _own_run(p, o, prs.BACKEND_FUSED, None, sort_every_step)
    for key in ("pos", "rad", "dead", "phase"):
        assert np.array_equal(_bits(own[0][key]), _bits(ref[0][key])), ("after reset()", key)
    for k in STEPS:
        for key in ("dead", "hash", "index"):
            assert np.array_equal(own[k][key], ref[k][key]), (k, key)
        vs = max(float(np.abs(ref[k]["vel"]).max()), 1e-3)
        assert util.rel_err(own[k]["pos"], ref[k]["pos"], 1.0) < 1e-5, k
        assert util.rel_err(own[k]["vel"], ref[k]["vel"], vs) < 1e-5, k
        assert util.rel_err(own[k]["rad"], ref[k]["rad"], 0.1) < 1e-5, k
        assert util.rel_err(own[k]["phase"], ref[k]["phase"], 1.0) < 1e-5, k


@pytest.mark.parametrize("name", util.CFGS)
@pytest.mark.parametrize("tag", ["refcadence", "sortall"])
def test_committed_goldens_equal_the_reference_class(name, tag, collide_kernel_default):
    """tests/golden/*.npz were produced by the reference's kernels under THIS repository's host logic
    (tests/golden/make_golden.py); the reference's own class must reproduce them bit for bit — which pins the CPU oracle,
    checked against the same files in tests/test_cpu_oracle.py, to the reference's host code as well"""
    _need_refhost()
    g = np.load(os.path.join(util.GOLDEN, f"{name}.{tag}.npz"))
    p, o, ref = _reference_run(name, tag == "sortall")
    assert np.array_equal(_bits(g["pos0"]), _bits(ref[0]["pos"])) and np.array_equal(_bits(g["rad0"]), _bits(ref[0]["rad"]))
    for k in STEPS:
        for key, rk in (("pos", "pos"), ("vel", "vel"), ("rad", "rad"), ("phase", "phase"), ("dead", "dead"), ("hash", "hash"),
                        ("index", "index"), ("fr", "absForce_r"), ("fa", "absForce_a")):
            assert np.array_equal(_bits(g[f"{key}_{k}"]), _bits(ref[k][rk])), (k, key)
        occ = g[f"occ_{k}"]
        assert np.array_equal(np.nonzero(ref[k]["cellStart"] != 0xFFFFFFFF)[0].astype(np.uint32), occ), k
        assert np.array_equal(ref[k]["cellStart"][occ], g[f"cs_occ_{k}"]) and np.array_equal(ref[k]["cellEnd"][occ], g[f"ce_occ_{k}"]), k


@pytest.mark.parametrize("name", util.CFGS)
@pytest.mark.parametrize("sort_every_step", [False, True])
@pytest.mark.parametrize("gl_build", [False, True])
def test_reference_class_linked_against_this_library(name, sort_every_step, gl_build, collide_kernel_default):
    """THE DROP-IN, EXECUTED: oracle/_ref/libprs_dropin.so is the reference's own
    particlebot.cpp, verbatim, linked against libparticlebot_b200.so instead of the reference's particlebot_cuda.o — every
    allocateArray / setParameters / curand_setup / updatePhase / updateRad_light_wave / integrateSystem / calcHash /
    sortParticlebots / reorderDataAndFindCellStart / collide / calcCOG / updateCol call of the class lands in this library.
    Same placement, same 100 steps, bit for bit, as the class over its own kernels.
    gl_build: the same class over the product's OPENGL BUILD (-DPRS_WITH_GL, oracle/_ref/libparticlebot_b200_glstub.so): the
    register / map / unmapGLBufferObject the class calls every step for its position, radius and colour VBOs
    (particlebot.cpp:105-113, 200-205) are the PRODUCT's, which go through the CUDA graphics API — redirected to the headless
    buffer objects here, there being no OpenGL on these boxes."""
    _need_refhost()
    path = ob.DROPIN_GL_PATH if gl_build else ob.DROPIN_PATH
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} not built (needs /root/reference at build time)")
    prs.lib().prs_set_stream(None)
    prs.lib().prs_set_world_half_extent(64.0)
    p, o, ref = _reference_run(name, sort_every_step)
    _, _, own = _reference_run(name, sort_every_step, path)
    for key in ("pos", "rad", "dead", "phase"):
        assert np.array_equal(_bits(own[0][key]), _bits(ref[0][key])), ("after reset()", key)
    for k in STEPS:
        occ = np.nonzero(ref[k]["cellStart"] != 0xFFFFFFFF)[0]
        assert np.array_equal(np.nonzero(own[k]["cellStart"] != 0xFFFFFFFF)[0], occ), k
        assert np.array_equal(own[k]["cellStart"][occ], ref[k]["cellStart"][occ]) and np.array_equal(own[k]["cellEnd"][occ], ref[k]["cellEnd"][occ]), k
        for key in ("dead", "hash", "index", "pos", "vel", "rad", "phase", "absForce_r", "absForce_a"):
            assert np.array_equal(_bits(own[k][key]), _bits(ref[k][key])), (k, key)
