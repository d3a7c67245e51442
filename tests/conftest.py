import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")


def load_module(name, path):
    """import a build script by path (both build scripts are called build.py)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without CUDA or without the built library."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    lib = os.path.join(ROOT, "pympc_quadruped_b200", "csrc", "_build", "libmpcq.so")
    if have and os.path.exists(lib):
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libmpcq.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY host emulation of the device code (see tests/emu/mpcq_emu.cpp)."""
    import ctypes
    return ctypes.CDLL(load_module("mpcq_emu_build", os.path.join(ROOT, "tests", "emu", "build.py")).build())


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
