import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The float32 CPU oracle (restated types) — the checker, never the thing under test on GPU runs."""
    from oracle.binding import Oracle, build
    build(ref=True)
    return Oracle(reference_types=False)


@pytest.fixture(scope="session")
def oracle_ref():
    """The same oracle built on the reference's own zenyth::math types (oracle/_ref)."""
    from oracle.binding import LIB_REFERENCE, Oracle, build
    build(ref=True)
    if not os.path.exists(LIB_REFERENCE):
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt .so)")
    return Oracle(reference_types=True)


@pytest.fixture(scope="session")
def zn_lib():
    """The product C-ABI library; built here if stale (nvcc cross-compiles without a GPU)."""
    from zenyth_b200 import build as zb
    zb.build()
    from zenyth_b200 import _capi
    return _capi.load()


@pytest.fixture(scope="session")
def gpu_ctx(zn_lib):
    import zenyth_b200 as zn
    ctx = zn.GpuContext(0)
    yield ctx
    ctx.Close()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_values(a, b):
    """Numerically identical (treats +0 and -0 as equal; NaN never equal)."""
    return bool(np.all(np.asarray(a) == np.asarray(b)))


def rel_err(a, b):
    """SURVEY §8d gate: max|a-b| / max(||b||_inf, 1e-30)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
