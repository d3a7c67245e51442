import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY host emulation of the device code (see tests/emu/mpcq_emu.cpp)."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build as emu_build
    return ctypes.CDLL(emu_build.build())


@pytest.fixture(scope="session", autouse=True)
def _single_threaded_blas():
    """The golden fixtures were generated with single-threaded BLAS; bitwise comparisons of the
    reference's big float64 matmuls need the same summation order."""
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        yield
        return
    with threadpool_limits(limits=1):
        yield
