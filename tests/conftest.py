import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    """without importing torch: ask the CUDA driver library directly"""
    import ctypes
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """-m gpu tests on a box without a CUDA device are skipped, not errors (the product has no CPU
    path; on the B200 box they all run)"""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the CUDA path has no CPU fallback")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def qg():
    import _pkg
    return _pkg.load()


@pytest.fixture(scope="session")
def pyorc():
    import pyorc
    pyorc.build()
    return pyorc
