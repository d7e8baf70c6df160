import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device and the built library: skip (not fail) where either is missing, so that a plain
    `pytest tests` on a CPU-only host is green and real regressions stay visible"""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    have_lib = os.path.exists(os.path.join(ROOT, "thinshelllab_b200", "libtsl.so"))
    if have_gpu and have_lib:
        return
    why = "no CUDA device" if not have_gpu else "thinshelllab_b200/libtsl.so not built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
