import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "proland-4.0_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
