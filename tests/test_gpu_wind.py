"""freq_grid_pathFromPath on the device: with winds the level grids are fac[ip] * freq_grid
(src/m_ppvar.cc:47-77, wind_shift src/m_frequency_grid.cc:4-55); the library applies the factor inside the
kernels instead of taking np shifted copies of the grid."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.conftest import assert_propmat_close

pytestmark = pytest.mark.gpu
C0 = 299792458.0


def _windy(c, scale=60.0, seed=0):
    rng = np.random.default_rng(seed)
    c.atm.wind = rng.normal(0.0, scale, (c.np_, 3))
    c.atm.los = np.tile(np.array([150.0, 40.0]), (c.np_, 1))
    return c


def _factors(c):
    """Direct evaluation: fac = 1 - wind . n / c with n the propagation direction (mirrored los)."""
    za, aa = np.deg2rad(180 - c.atm.los[:, 0]), np.deg2rad(c.atm.los[:, 1] + 180)
    n = np.stack([np.sin(za) * np.sin(aa), np.sin(za) * np.cos(aa), np.cos(za)], axis=1)  # u (east), v (north), w (up)
    return 1.0 - (c.atm.wind * n).sum(axis=1) / C0


def test_wind_equals_explicitly_shifted_grids(wsm, orc):
    c = _windy(synth.tiny_case(nl=64, nf=300, np_=6))
    fac = _factors(c)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    Kr, _ = orc.propmat_levels(c.cat, c.f, c.atm)
    assert_propmat_close(K, Kr)
    # the same thing the long way round: np shifted copies of the grid, no wind
    import copy

    calm = copy.deepcopy(c.atm)
    calm.wind = None
    f2 = fac[:, None] * c.f[None, :]
    K2, _ = wsm.spectral_propmat_pathFromPath(c.cat, f2, calm)
    assert_propmat_close(K, K2, rtol=1e-6)  # fac from the closed form above agrees with wind_shift to ~1e-16 in f
    I, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    Ir, _ = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, c.f), orc.planck_tb(c.f, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6
    # winds matter in this fixture (Doppler shift ~ 1e-7 f moves the line centres)
    I0, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, calm, c.r, c.I_bkg)
    assert np.abs(I - I0).max() > 0


def test_zero_wind_is_bitwise_no_wind(wsm):
    c = synth.tiny_case(nl=64, nf=300, np_=5)
    I0, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    c.atm.wind = np.zeros((c.np_, 3))
    I1, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    assert np.array_equal(I0, I1)


def test_wind_with_zeeman_cutoff_and_jacobians(wsm, orc):
    c = _windy(synth.case_c3(nf=38 * 8, np_=5, los=(150.0, 40.0)), scale=30.0, seed=2)
    tg = (("T",), ("VMR", 0))
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, targets=tg, hse_derivative=1)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1)
    np.testing.assert_allclose(I, Ir, rtol=1e-9, atol=1e-12 * np.abs(Ir).max())
    for q in range(2):
        sc = np.abs(dIr[:, :, q]).reshape(-1, 4).max(axis=0)
        assert (np.abs(dI[:, :, q] - dIr[:, :, q]).reshape(-1, 4).max(axis=0) <= 5e-7 * sc + 1e-300).all()
    c2 = _windy(synth.case_c1(nl=80, nf=400, cutoff=1.5e9), scale=200.0, seed=3)
    Kr, _ = orc.propmat_levels(c2.cat, c2.f, c2.atm)
    K, _ = wsm.spectral_propmat_pathFromPath(c2.cat, c2.f, c2.atm)
    assert_propmat_close(K, Kr, atol_scale=1e-11)


def test_unphysical_wind_is_rejected(wsm):
    c = synth.tiny_case(nl=16, nf=32, np_=2)
    c.atm.los = np.tile(np.array([180.0, 0.0]), (2, 1))
    c.atm.wind = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 4e8]])  # faster than light along the propagation direction
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    assert e.value.code == abi.ERR_INVALID and "frequency scaling" in str(e.value)
