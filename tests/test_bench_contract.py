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
