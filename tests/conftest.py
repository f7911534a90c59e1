import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present():
    """a CUDA driver with at least one device (no torch import: this runs at collection time)"""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cu.cuInit(0) == 0 and cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """without a GPU the tests marked `gpu` are skipped, not failed: `pytest tests` is then the CPU suite.  (The product has no
    CPU fallback: what is skipped is the test, nothing runs elsewhere.)"""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def plb():
    """The ctypes binding of the CUDA library (built on demand; no fallback)."""
    import proland_b200
    if not os.path.exists(proland_b200.LIB_PATH):
        proland_b200.build()
    return proland_b200


@pytest.fixture()
def ctx(plb):
    c = plb.Context(0)
    yield c
    c.close()
