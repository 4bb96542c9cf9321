"""GPU parity of the register-resident w(z) against the reference's Faddeeva::w.

Reference call site: src/core/lbl/lbl_lineshape_voigt_lte.cpp:239; known-answer vectors:
3rdparty/Faddeeva/Faddeeva.cc:4041-4195 (tests/golden/faddeeva_kat.json).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "faddeeva_kat.json")


def _relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_kat_57_points(wsm):
    d = json.load(open(GOLD))
    z = np.array([complex(float(a), float(b)) for a, b in d["z"]])
    w = np.array([complex(float(a), float(b)) for a, b in d["w"]])
    ok = np.isfinite(z.real) & np.isfinite(z.imag) & (np.abs(z) < 1e100)
    got = wsm.faddeeva_w(z[ok])
    ref = w[ok]
    # the reference's own criterion: relative error per part <= 1e-13 (Faddeeva.cc:4222)
    re = np.where(ref.real == 0, np.abs(got.real), _relerr(got.real, ref.real))
    im = np.where(ref.imag == 0, np.abs(got.imag), _relerr(got.imag, ref.imag))
    assert ok.sum() >= 40
    assert re.max() <= 1e-13, (z[ok][np.argmax(re)], re.max())
    assert im.max() <= 1e-13, (z[ok][np.argmax(im)], im.max())


@pytest.mark.parametrize("seed", [0, 1])
def test_dense_sweep_vs_reference(wsm, orc, seed):
    rng = np.random.default_rng(seed)
    n = 200_000
    x = np.concatenate([rng.uniform(-30, 30, n), 10 ** rng.uniform(-8, 7.5, n) * rng.choice([-1, 1], n)])
    y = np.concatenate([10 ** rng.uniform(-14, 1, n), 10 ** rng.uniform(-6, 7.5, n)])
    y[:500] = 0.0
    z = x + 1j * y
    got = wsm.faddeeva_w(z)
    ref = orc.faddeeva_w(z)
    # complex relative error; the reference package itself is accurate to ~1e-13
    assert _relerr(got, ref).max() <= 2e-13
    big = np.abs(ref.real) > 1e-290
    assert _relerr(got.real[big], ref.real[big]).max() <= 2e-13
    # Im w crosses zero at x = 0: relative to |w|
    assert (np.abs(got.imag - ref.imag) / np.abs(ref)).max() <= 2e-13
    # symmetry w(-x + iy) = conj(w(x + iy)) must hold bit for bit
    gm = wsm.faddeeva_w(-z.real + 1j * z.imag)
    assert np.array_equal(gm.real, got.real) and np.array_equal(gm.imag, -got.imag)


def test_complex_dawson_against_the_reference_object(wsm, orc):
    """Faddeeva::Dawson(z) for complex z (Faddeeva.cc:461-570), the element-wise function of rtepack::dawson(specmat) in
    polarised linprop layers: every branch of the package (both axes, the |z| < 5e-3 Taylor series, the expansion about the
    real axis for small |y| and |xy| incl. |x| > 40 and |x| > 5e7, both half planes of the general formula)."""
    rng = np.random.default_rng(5)
    z = [0.0, 1e-3, -2.5, 30.0, 1e-3j, -1e-3j, 0.3j, -2j, 4j, 1e-3 + 1e-3j, -4e-3 + 2e-3j, 2.0 + 1e-3j, 0.7 - 4e-3j,
         60.0 + 5e-5j, -45.0 - 1e-5j, 6e7 + 1e-11j, 1.5 + 0.5j, -1.5 - 0.5j, 8.0 + 0.2j, 3.0 - 1.0j, 0.01 + 0.01j]
    mag = 10 ** rng.uniform(-3, 1.2, 4000)
    ang = rng.uniform(-np.pi, np.pi, 4000)
    z = np.concatenate([np.array(z, dtype=np.complex128), mag * np.exp(1j * ang), rng.uniform(-6, 6, 2000) + 1j * 10 ** rng.uniform(-6, -2, 2000)])
    z = z[np.abs(z.imag) < np.abs(z.real) + 3.0]  # exp(y^2 - x^2) stays in range
    got, ref = wsm.dawson(z), orc.dawson(z)
    assert np.all(np.isfinite(ref)) and np.all(np.isfinite(got))
    err = np.abs(got - ref) / np.abs(np.where(ref == 0, 1.0, ref))
    assert err.max() <= 2e-12, (err.max(), z[np.argmax(err)])
