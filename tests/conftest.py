import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    """True when libolapgpu.so can bind a CUDA device (olap_init(0)); no torch import needed."""
    try:
        from olap_in_memory_b200 import _native

        return _native.lib().olap_init(0) == 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a GPU skips the gpu-marked tests instead of failing
    them one by one (the store has no CPU fallback: every one of them would raise OlapError)."""
    if not any("gpu" in item.keywords for item in items) or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the store has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
