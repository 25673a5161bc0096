import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): C restatement of the reference path."""
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def vg():
    """The product: ctypes view of libvisgeom_b200.so (built on demand; CUDA only)."""
    from visgeom_b200 import build as _b
    _b.build()
    import visgeom_b200
    return visgeom_b200


@pytest.fixture(scope="session")
def gpu(vg):
    if vg.device_count() < 1:
        pytest.fail("gpu-marked test started without a CUDA device: the engine has no CPU fallback")
    return vg
