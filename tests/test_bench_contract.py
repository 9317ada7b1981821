"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port, the one arm that runs on the
host cores) prints ONE JSON line with the keys the driver reads, and the native arm refuses to run without a device
instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

from tests import util

BENCH = os.path.join(util.ROOT, "bench.py")


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "2", "--warmup", "1", "--robots-log2", "14"],
                       cwd=util.ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_native_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "0", "--robots-log2", "12", "--no-cpu-baseline",
                        "--no-ref-cuda"], cwd=util.ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0                      # no CPU fallback: the product path needs the CUDA device
    assert not [l for l in r.stdout.splitlines() if l.startswith('{"metric"')]


def test_reference_arm_does_not_map_the_product_library():
    """the CPU arm builds its parameters with oracle/params.py: libparticlebot_b200.so must not be loaded by it"""
    code = ("import sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--robots-log2', '12']\n"
            "import bench; bench.main()\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libprs_oracle' in maps\n"
            "assert 'libparticlebot_b200' not in maps, 'product library mapped by the reference arm'\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=util.ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]


def test_oracle_params_restatement_equals_the_product_parser():
    """oracle/params.py (pure Python, used by the reference arm) against prs_config.cpp on the shipped cfgs without
    obstacle lists, field by field, plus the synthetic worlds of bench.swarm_config"""
    import ctypes as C
    import particlerobotsimulations_b200 as prs
    from oracle import params as op
    import bench
    skip = {"x1obs", "x2obs", "y1obs", "y2obs", "x_cir_obs", "y_cir_obs", "r_cir_obs", "_pad0", "seed"}

    def same(a, b, seed_too):
        for name, typ in prs.SimParams._fields_:
            if name in skip and not (name == "seed" and seed_too):
                continue
            va, vb = getattr(a, name), getattr(b, name)
            if hasattr(va, "x"):
                assert (va.x, va.y) == (vb.x, vb.y), name
            else:
                assert va == vb, (name, va, vb)

    pa, _ = op.defaults()
    pb, ob_ = prs.default_params()
    same(pa, pb, False)          # the default seed is time(NULL) in the reference
    for name in ("example", "example_dead_cells", "example_object_transport"):
        pa, ra = op.load_cfg(os.path.join(util.EXAMPLES, name + ".cfg"))
        pb, rb = util.cfg(name)
        same(pa, pb, True)
        assert abs(ra["timestep"] - rb.timestep) == 0 and ra["sort_interval"] == rb.sort_interval
    for log2n in (14, 20, 26):
        pa, oa, ga = bench.swarm_config(None, log2n)
        pb, ob2, gb = bench.swarm_config(prs, log2n)
        same(pa, pb, True)
        assert ga == gb and oa.timestep == ob2.timestep
