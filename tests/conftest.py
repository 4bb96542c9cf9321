import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/_ref/liboracle.so): the checker, never the thing under test on the GPU side."""
    from tests import oracle_lib

    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def wsm():
    """The product's host mirror over the C ABI; fails loudly when the CUDA library is missing."""
    from arts_b200 import wsm as _wsm

    _wsm.lib()
    return _wsm


def assert_propmat_close(K, Kref, rtol=1e-9, atol_scale=1e-12, what="propmat"):
    """|dK| <= rtol |Kref| + atol_scale * max|Kref[..., c]| per Propmat component c.

    north_star: relative error <= 1e-9 on propmat elements; the per-component absolute
    floor only covers the zero crossings of the dispersive / polarised components.
    """
    K = np.asarray(K)
    Kref = np.asarray(Kref)
    assert K.shape == Kref.shape, (K.shape, Kref.shape)
    assert np.isfinite(K).all(), f"{what}: non-finite values"
    scale = np.abs(Kref).reshape(-1, Kref.shape[-1]).max(axis=0)
    tol = rtol * np.abs(Kref) + atol_scale * scale
    err = np.abs(K - Kref)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err / np.maximum(tol, 1e-300)), err.shape)
        raise AssertionError(
            f"{what}: {bad.sum()} of {bad.size} elements off; worst at {i}: got {K[i]!r} ref {Kref[i]!r} "
            f"rel {err[i] / max(abs(Kref[i]), 1e-300):.3e}"
        )
