import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        import ctypes

        cuda = ctypes.CDLL("libcuda.so.1")
        if cuda.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import PortOracle, build

    build()
    return PortOracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import RefOracle, build, have_ref

    build()
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return RefOracle()


@pytest.fixture(scope="session")
def golden0():
    from tools import frames as F

    return F.load_golden("kitti_f000")


@pytest.fixture(scope="session")
def golden100():
    from tools import frames as F

    return F.load_golden("kitti_f100")
