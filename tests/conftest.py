import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Product library (nvcc, sm_100a; cross-compiles without a GPU) and the CPU oracle."""
    from particlerobotsimulations_b200 import build as prs_build
    prs_build.build()
    from oracle import binding
    if not os.path.exists(binding.LIB_PATH):
        binding.build()
    return True


@pytest.fixture
def collide_kernel_default():
    """library defaults for the knobs other test modules vary per test"""
    import particlerobotsimulations_b200 as prs
    L = prs.lib()
    L.prs_set_collide_warp_max(16384)
    L.prs_set_collide_tile(0)
    L.prs_set_pdl(1)
    L.prs_set_k1_x2(1)
    L.prs_set_collide_dense(1)
    L.prs_set_fuse_gather_max(65536)
    L.prs_bin_set_mode(0)
    yield
