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


# ---------------------------------------------------------------------------------------------------------------------
# Wind Jacobians: d propmat / d wind = (dscl shape + scl sum s inv_gd dF)(f) * f * freq_wind_shift_jac
# (single_shape::df lbl_lineshape_voigt_lte.cpp:275, df_core_calc :1036-1062, compute_derivative :1514-1523,
# spectral_propmat_jacWindFix src/m_frequency_grid.cc:106-182), and on through the Stokes chain like any other target
# ---------------------------------------------------------------------------------------------------------------------
from tests.test_gpu_jacobian import assert_jac_close  # noqa: E402

WIND = (("wind_u",), ("wind_v",), ("wind_w",))


def test_wind_jacobian_propmat_scalar_and_mixed_targets(wsm, orc):
    c = _windy(synth.tiny_case(nl=200, nf=600, np_=5), scale=40.0, seed=4)
    tg = (("wind_u",), ("T",), ("wind_w",), ("VMR", 0), ("wind_v",))  # two passes of the Jacobian kernel
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    assert_propmat_close(K, Kr)
    assert np.abs(dKr[:, 0]).max() > 0 and np.abs(dKr[:, 2]).max() > 0
    for q in range(len(tg)):
        for lev in range(c.np_):
            assert_jac_close(dK[lev, q], dKr[lev, q], what=f"dK target {tg[q]} level {lev}")


@pytest.mark.parametrize("cutoff", [None, 1.5e9])
def test_wind_jacobian_far_and_near_tiles(wsm, orc, cutoff):
    """Wide grid over many tiles: closed-form far tiles, per-pair far path, near pairs; with and without ByLine cutoff."""
    c = _windy(synth.case_c1(nl=700, nf=1500, cutoff=cutoff), scale=100.0, seed=6)
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=WIND)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=WIND)
    assert_propmat_close(K, Kr, atol_scale=1e-11)
    for q in range(3):
        assert_jac_close(dK[:, q], dKr[:, q], what=f"dK wind target {q}")


def test_wind_jacobian_zeeman_mixing_through_the_fused_chain(wsm, orc):
    c = _windy(synth.case_c3(nf=38 * 8, np_=5, los=(150.0, 40.0)), scale=30.0, seed=2)
    tg = (("wind_v",), ("T",), ("wind_w",))
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    for q in range(3):
        assert_jac_close(dK[:, q], dKr[:, q], rtol=5e-7, what=f"Zeeman dK target {tg[q]}")
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, targets=tg, hse_derivative=1)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1)
    np.testing.assert_allclose(I, Ir, rtol=1e-9, atol=1e-12 * np.abs(Ir).max())
    assert np.abs(dIr[:, :, 0]).max() > 0
    for q in range(3):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=5e-7, what=f"dI target {tg[q]}")


def test_wind_jacobian_of_a_calm_path(wsm, orc):
    """No winds at all: the reference still runs wind_shift at every point, and its zero-wind special cases give all
    three components -cos(za_p) / c (src/m_frequency_grid.cc:56-80: df = 1, angle derivatives 0).  Same rows with a
    wind array of zeros and with none."""
    c = synth.tiny_case(nl=64, nf=300, np_=4)
    c.atm.los = np.tile(np.array([130.0, -20.0]), (c.np_, 1))
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, targets=WIND)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=WIND)
    for q in range(3):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=5e-7, what=f"calm dI target {q}")
    assert np.abs(dI[:, :, 0]).max() > 0
    assert np.array_equal(dI[:, :, 0], dI[:, :, 1]) and np.array_equal(dI[:, :, 0], dI[:, :, 2])
    c.atm.wind = np.zeros((c.np_, 3))
    I2, dI2 = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=WIND)
    assert np.array_equal(I, I2) and np.array_equal(dI, dI2)


def test_wind_jacobian_mirrored_band(wsm, orc):
    c = _windy(synth.tiny_case(nl=200, nf=500, np_=3), scale=50.0, seed=8)
    rng = np.random.default_rng(21)
    for b in range(len(c.cat.band_isot)):
        lo, hi = c.cat.band_offset[b], c.cat.band_offset[b + 1]
        c.cat.f0[lo:hi] = np.sort(rng.uniform(1e9, 60e9, hi - lo))
    c.f = np.linspace(0.5e9, 80e9, c.nf)
    c.cat.band_lineshape[:] = abi.LINESHAPE_VP_LTE_MIRROR
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=WIND[:2])
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=WIND[:2])
    assert_propmat_close(K, Kr)
    for q in range(2):
        assert_jac_close(dK[:, q], dKr[:, q], what=f"mirrored dK wind target {q}")


def test_addlines_leaves_the_frequency_derivative_for_the_agendas_wind_fix(wsm, orc):
    """Per-level agenda flow of the reference: freq_gridWindShift, spectral_propmatAddLines (rows = d/df), then
    spectral_propmat_jacWindFix (x * f * freq_wind_shift_jac, src/m_frequency_grid.cc:106-182) - equal to the one-call path."""
    import copy

    c = _windy(synth.tiny_case(nl=100, nf=300, np_=3), scale=40.0, seed=9)
    tg = (("wind_u",), ("T",), ("wind_w",))
    _, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    for ip in range(c.np_):
        fac, jac = orc.wind_shift(c.atm.wind[ip], c.atm.los[ip])
        fs = fac * c.f  # freq_gridWindShift
        pt = c.atm.take([ip])
        pt.wind = None  # the grid is already shifted
        k, dk = np.zeros((c.nf, 7)), np.zeros((3, c.nf, 7))
        wsm.spectral_propmatAddLines(k, dk, fs, tg, abi.SPECIES_BATH, c.cat, pt)
        dk[0] *= (fs * jac[0])[:, None]
        dk[2] *= (fs * jac[2])[:, None]
        for q in range(3):
            assert_jac_close(dk[q], dK[ip, q], rtol=1e-12, what=f"level {ip} target {tg[q]}")
