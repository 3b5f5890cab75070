import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _alias_lib_module():
    """tests import the ctypes binding as `tensorbranching_lib.L` (the package directory has a dot in its name)"""
    import types

    import tbcuda  # noqa: F401
    mod = types.ModuleType("tensorbranching_lib")
    mod.L = sys.modules["tbcuda._lib"]
    sys.modules["tensorbranching_lib"] = mod


_alias_lib_module()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tb():
    import tbcuda

    return tbcuda


def _cuda_device_present():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu need a CUDA device and the built engine: without either they are SKIPPED (so a CPU-only box
    reports skips, not errors).  On the GPU box nothing is skipped: a missing libtbcuda.so fails loudly there."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this host (gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def engine(tb):
    eng = tb.Engine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine_dataflow(tb):
    """the dataflow executor forced for every call (TB_DATAFLOW=1); the default engine picks per call"""
    os.environ["TB_DATAFLOW"] = "1"
    try:
        eng = tb.Engine(0)
    finally:
        del os.environ["TB_DATAFLOW"]
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine_levelsync(tb):
    """the round-1 executor (one launch per dependency level and kernel kind), kept for A/B testing: TB_LEVEL_SYNC=1"""
    os.environ["TB_LEVEL_SYNC"] = "1"
    try:
        eng = tb.Engine(0)
    finally:
        del os.environ["TB_LEVEL_SYNC"]
    yield eng
    eng.close()
