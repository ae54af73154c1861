import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _ensure_library():
    """The CPU suite checks that the C-ABI library loads: build it (nvcc cross-compiles without a GPU) if missing."""
    lib = os.path.join(ROOT, "gym_ballenv_b200", "libballenv_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


_ensure_library()
