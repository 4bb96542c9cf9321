"""GPU parity of stage 1 (spectral_propmatAddLines / spectral_propmat_pathFromPath) through the C ABI
against the CPU oracle (engine A of the reference, src/core/lbl/lbl_lineshape_voigt_lte.cpp:1652-1725)."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.conftest import assert_propmat_close

pytestmark = pytest.mark.gpu


def test_c1_full_config(wsm, orc):
    """BASELINE config 1: 1 species, 1k Voigt lines, 1e4 frequencies, 1 level; <= 1e-9 relative."""
    c = synth.case_c1()
    Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm)
    K = np.zeros((c.nf, 7))
    wsm.spectral_propmatAddLines(K, None, c.f, (), abi.SPECIES_BATH, c.cat, c.atm)
    assert_propmat_close(K[None], Kref)
    assert (K[:, 0] > 0).all() and not K[:, 1:].any()


def test_addlines_accumulates(wsm, orc):
    """+= semantics of lbl_lineshape_voigt_lte.cpp:1691 on the caller's array."""
    c = synth.case_c1(nl=100, nf=777)
    Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm)
    K = np.full((c.nf, 7), 0.25)
    wsm.spectral_propmatAddLines(K, None, c.f, (), abi.SPECIES_BATH, c.cat, c.atm)
    assert_propmat_close((K - 0.25)[None], Kref, rtol=1e-9, atol_scale=1e-9)
    assert np.array_equal(K[:, 1:], np.full((c.nf, 6), 0.25))


@pytest.mark.parametrize("nf,nl,np_", [(1, 2, 1), (257, 64, 6), (513, 600, 3), (1025, 38, 2)])
def test_ragged_sizes(wsm, orc, nf, nl, np_):
    c = synth.tiny_case(nl=nl, nf=nf, np_=np_)
    Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    assert_propmat_close(K, Kref)


def test_select_species_and_sum(wsm, orc):
    c = synth.tiny_case(nl=200, nf=400, np_=4)
    cat = wsm.Catalog(c.cat)
    Kall, _ = wsm.spectral_propmat_pathFromPath(cat, c.f, c.atm)
    parts = []
    for s in range(c.cat.n_species):
        Ks, _ = wsm.spectral_propmat_pathFromPath(cat, c.f, c.atm, select_species=s)
        Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm, select_species=s)
        assert_propmat_close(Ks, Kref)
        parts.append(Ks)
    assert_propmat_close(sum(parts), Kall, rtol=1e-13)
    cat.close()


def test_per_level_frequency_grids(wsm, orc):
    """freq_grid_path: one (wind-shifted) grid per level (src/m_ppvar.cc:47-77)."""
    c = synth.tiny_case(nl=100, nf=300, np_=5)
    f2 = c.f[None, :] * (1 + 1e-6 * np.arange(c.np_)[:, None])
    Kref, _ = orc.propmat_levels(c.cat, f2, c.atm)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, f2, c.atm)
    assert_propmat_close(K, Kref)


@pytest.mark.parametrize("cutoff", [0.3e9, 2e9, 750e9])
def test_byline_cutoff(wsm, orc, cutoff):
    """LineByLineCutoffType::ByLine (lbl_lineshape_voigt_lte.cpp:591-616, lbl_data.cpp:61-68)."""
    c = synth.case_c1(nl=300, nf=1500, cutoff=cutoff)
    Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm, no_negative_absorption=0)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, no_negative_absorption=0)
    assert_propmat_close(K, Kref, atol_scale=1e-11)
    Kref1, _ = orc.propmat_levels(c.cat, c.f, c.atm, no_negative_absorption=1)
    K1, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, no_negative_absorption=1)
    assert_propmat_close(K1, Kref1, atol_scale=1e-11)


@pytest.mark.parametrize("cutoffs", [(3e9,), (1.5e9, None, 6e9), (0.2e9, 40e9)])
def test_byline_cutoff_many_tiles(wsm, orc, cutoffs):
    """Real bands with ByLine cutoffs are merged per species and summed by the real kernel with one window per line:
    enough lines and frequencies that (block, tile) pairs outside every window, inside every window and cut by
    window edges all occur, with bands of different cutoffs (and none) interleaved in one species."""
    nb = len(cutoffs)
    per_band = 4000
    rng = np.random.default_rng(11)
    c = synth.case_c1(nl=nb * per_band, nf=6000, seed=3)
    for b in range(nb):  # each band sorted by f0 (lbl_data.cpp:61-68), bands interleaved in frequency
        c.cat.f0[b * per_band:(b + 1) * per_band] = np.sort(rng.uniform(100e9, 130e9, per_band))
    c.cat.band_isot = np.zeros(nb, np.int32)
    c.cat.band_lineshape = np.zeros(nb, np.int32)
    c.cat.band_offset = np.arange(nb + 1, dtype=np.int64) * per_band
    c.cat.band_cutoff_type = np.array([abi.CUTOFF_NONE if v is None else abi.CUTOFF_BYLINE for v in cutoffs], np.int32)
    c.cat.band_cutoff_value = np.array([np.inf if v is None else v for v in cutoffs])
    cat = wsm.Catalog(c.cat)
    assert cat.counts()[0] == nb * per_band, "all bands must land in the merged real segment"
    for clamp in (0, 1):
        Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm, no_negative_absorption=clamp)
        K, _ = wsm.spectral_propmat_pathFromPath(cat, c.f, c.atm, no_negative_absorption=clamp)
        assert_propmat_close(K, Kref, atol_scale=1e-11)
    # a sub-grid sees other window bounds of active_lines but the same values (shard invariance with cutoffs)
    fs = c.f[1000:2500]
    path = wsm.Path(cat, len(fs), 1)
    path.set_grid_bounds([c.f[0], c.f[-1]])  # before upload: active_lines sees the whole grid
    path.upload(fs, c.atm, np.zeros(0), np.zeros((len(fs), 4)), no_negative_absorption=1)
    path.run_propmat()
    Ks = np.empty((1, len(fs), 7))
    path.download(K=Ks)
    assert np.array_equal(Ks, K[:, 1000:2500])
    path.close()
    cat.close()


@pytest.mark.parametrize("los", [(180.0, 0.0), (120.0, 30.0)])
def test_zeeman_polarised(wsm, orc, los):
    """BASELINE config 3 (reduced nf): full 7-component Propmat with Zeeman sub-lines and line mixing."""
    c = synth.case_c3(nf=38 * 40, np_=7, los=los)
    Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm)
    cat = wsm.Catalog(c.cat)
    assert sum(cat.counts()[1:]) == 4332 and cat.counts()[0] == 0  # 4446 components, 114 with zero strength popped (:354-357)
    K, _ = wsm.spectral_propmat_pathFromPath(cat, c.f, c.atm)
    assert_propmat_close(K, Kref)
    assert np.abs(K[..., 1:]).max() > 0
    cat.close()


def test_line_mixing_clamp(wsm, orc):
    """Per-band negative clamp (lbl_lineshape_voigt_lte.cpp:1688-1692) with strong line mixing."""
    c = synth.case_c1(nl=40, nf=900)
    c.cat.ls_type[:, abi.VAR_Y] = abi.TM_T1
    rng = np.random.default_rng(5)
    c.cat.ls_X[:, abi.VAR_Y, 0] = rng.uniform(-3e-3, 3e-3, len(c.cat.ls_species))
    c.cat.ls_X[:, abi.VAR_Y, 1] = 0.8
    c.cat.ls_type[:, abi.VAR_G] = abi.TM_T1
    c.cat.ls_X[:, abi.VAR_G, 0] = rng.uniform(-1e-9, 1e-9, len(c.cat.ls_species))
    c.cat.ls_X[:, abi.VAR_G, 1] = 0.5
    for clamp in (0, 1):
        Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm, no_negative_absorption=clamp)
        K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, no_negative_absorption=clamp)
        assert_propmat_close(K, Kref, atol_scale=1e-11)
    Kref0, _ = orc.propmat_levels(c.cat, c.f, c.atm, no_negative_absorption=0)
    assert (Kref0[..., 0] < 0).any(), "fixture must exercise the clamp"


def test_temperature_models(wsm, orc):
    """Every LineShapeModelType (lbl_temperature_model.h:62-283) through the prepare kernel."""
    c = synth.case_c1(nl=90, nf=400)
    n = len(c.cat.ls_species)
    rng = np.random.default_rng(7)
    types = [abi.TM_T0, abi.TM_T1, abi.TM_T2, abi.TM_T3, abi.TM_T4, abi.TM_T5, abi.TM_AER, abi.TM_DPL, abi.TM_POLY]
    t = np.array(types)[np.arange(n) % len(types)]
    c.cat.ls_type[:, abi.VAR_G0] = t
    X = c.cat.ls_X[:, abi.VAR_G0]
    X[:, 0] = rng.uniform(1e4, 3e4, n)
    X[:, 1] = rng.uniform(0.5, 1.0, n)
    X[:, 2] = rng.uniform(0.0, 0.3, n)
    X[:, 3] = rng.uniform(0.0, 1.0, n)
    aer = t == abi.TM_AER
    X[aer, 1:] = X[aer, :1] * rng.uniform(0.8, 1.2, (aer.sum(), 3))
    poly = t == abi.TM_POLY
    X[poly, 1] = 10.0
    X[poly, 2] = 1e-2
    X[poly, 3] = 1e-5
    t3 = t == abi.TM_T3
    X[t3, 1] = 20.0
    dpl = t == abi.TM_DPL
    X[dpl, 2] = X[dpl, 0] * 0.1
    for T in (230.0, 260.0, 300.0):
        c.atm.T[:] = T
        Kref, _ = orc.propmat_levels(c.cat, c.f, c.atm)
        K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
        assert_propmat_close(K, Kref)


def test_errors_are_loud(wsm):
    c = synth.tiny_case()
    c.cat.band_lineshape[0] = abi.LINESHAPE_OTHER
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.Catalog(c.cat)
    assert e.value.code == abi.ERR_UNSUPPORTED and "VP_LTE" in str(e.value)
    c = synth.tiny_case()
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, select_species=99)
    assert e.value.code == abi.ERR_INVALID
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=7)
    assert e.value.code == abi.ERR_INVALID
    with pytest.raises(ValueError):
        wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg[:-1])


def test_mid_wing_closed_form_accuracy(wsm, orc):
    """Where the reference still runs its continued fraction (1000 < |x| + y < 4000) the real line sum uses closed forms:
    the two-term far form from 1000 in the line-by-line kernel (2.5/x^4 <= 2.5e-12 relative), and the far-field sums
    (lbl_fmm.cu) the four-term form from 48 (<= 1.3e-12).  Both far inside
    the 1e-9 bound."""
    c = synth.case_c1(nl=1, nf=4000)
    c.atm.P[:] = 5.0  # nearly pure Doppler line: y << 1, x = (f - f0') / G_D
    f0 = c.cat.f0[0]
    gd = np.sqrt(2000 * 1.380649e-23 * 6.02214076e23 / 299792458.0**2 * 250.0 / 31.9898) * f0
    x = np.linspace(300.0, 9000.0, c.nf)
    c.f = f0 + x * gd
    Kr, _ = orc.propmat_levels(c.cat, c.f, c.atm)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    rel = np.abs(K[0, :, 0] - Kr[0, :, 0]) / Kr[0, :, 0]
    assert rel.max() <= 4e-12, (rel.max(), x[np.argmax(rel)])
    # the line-by-line kernel (AB200_FARFIELD=0)
    import os

    os.environ["AB200_FARFIELD"] = "0"
    try:
        K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    finally:
        del os.environ["AB200_FARFIELD"]
    rel = np.abs(K[0, :, 0] - Kr[0, :, 0]) / Kr[0, :, 0]
    assert rel.max() <= 4e-12, (rel.max(), x[np.argmax(rel)])
    assert rel[(x > 1000) & (x < 4000)].max() > 1e-15  # the two-term closed form is really in use there


def test_near_wing_closed_form_accuracy(wsm, orc):
    """Between |x| + y = 48 (MID_LIMIT) and the far limit the forward line sums evaluate the four-term continued fraction as
    one rational function instead of the reference's nu(z)-term recurrence: <= 1.3e-12 relative on Re w and Im w; below 48 the
    reference's own regions run.  One nearly Doppler line (y << x), one pressure-broadened line (y ~ x), and a line-mixing
    line through the complex kernel (both parts of w)."""
    for P, ymix in ((5.0, False), (3e4, False), (5.0, True)):
        c = synth.case_c1(nl=1, nf=3000)
        c.atm.P[:] = P
        if ymix:
            c.cat.ls_type[:, abi.VAR_Y] = abi.TM_T1
            c.cat.ls_X[:, abi.VAR_Y, 0] = 2e-3 / P
            c.cat.ls_X[:, abi.VAR_Y, 1] = 0.8
        f0 = c.cat.f0[0]
        gd = np.sqrt(2000 * 1.380649e-23 * 6.02214076e23 / 299792458.0**2 * 250.0 / 31.9898) * f0
        x = np.concatenate([-np.geomspace(900.0, 20.0, 1500), np.geomspace(20.0, 900.0, 1500)])
        c.f = np.sort(f0 + x * gd)
        Kr, _ = orc.propmat_levels(c.cat, c.f, c.atm, no_negative_absorption=0)
        K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, no_negative_absorption=0)
        rel = np.abs(K[0, :, 0] - Kr[0, :, 0]) / np.abs(Kr[0, :, 0])
        assert rel.max() <= 5e-12, (P, ymix, rel.max(), c.f[np.argmax(rel)])
        xs = np.abs(c.f - f0) / gd
        assert rel[(xs > 60) & (xs < 800)].max() > 1e-15, "the closed form is really in use there"
        assert rel[xs < 40].max() <= 2e-12  # the reference's own regions, through the ratio-form continued fraction


def test_small_grid_geometry_is_invisible(wsm):
    """Small problems run narrower frequency blocks (64 or 128 instead of 512 per CTA).  The value at a frequency must not
    depend on that: the same frequencies inside a grid large enough for the default geometry give the same bits."""
    c = synth.case_c1(nl=1000, nf=3000)
    K_small, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)             # 3000 frequencies: 64-wide blocks
    K_mid, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f[:40_000 // 16], c.atm)  # 2500: still small
    big = np.sort(np.concatenate([c.f, np.linspace(99e9, 131e9, 197_001)]))       # 200 001: default geometry
    K_big, _ = wsm.spectral_propmat_pathFromPath(c.cat, big, c.atm)
    idx = np.searchsorted(big, c.f)
    assert np.array_equal(big[idx], c.f)
    assert np.array_equal(K_big[:, idx], K_small)
    assert np.array_equal(K_mid, K_small[:, :2500])
    c2 = synth.case_c1(nl=1000, nf=30_000)                                        # 128-wide blocks
    K2, _ = wsm.spectral_propmat_pathFromPath(c2.cat, c2.f, c2.atm)
    big2 = np.sort(np.concatenate([c2.f, np.linspace(99e9, 131e9, 170_001)]))
    K2b, _ = wsm.spectral_propmat_pathFromPath(c2.cat, big2, c2.atm)
    assert np.array_equal(K2b[:, np.searchsorted(big2, c2.f)], K2)
