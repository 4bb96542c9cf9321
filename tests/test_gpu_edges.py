"""Edge cases of the path through the C ABI: empty and degenerate shapes, a single level, lines that
contribute nothing, extreme parameters — the cases the reference's shape checks and early returns cover
(m_lbl.cc:256-271, lbl_lineshape_voigt_lte.cpp:1663-1669, rtepack_transmission.cc:1268-1276)."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.conftest import assert_propmat_close

pytestmark = pytest.mark.gpu


def test_empty_frequency_grid(wsm):
    c = synth.tiny_case(nl=16, nf=8, np_=3)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f[:0], c.atm)
    assert K.shape == (3, 0, 7)
    I, _ = wsm.spectral_radClearskyEmission(c.cat, c.f[:0], c.atm, c.r, c.I_bkg[:0])
    assert I.shape == (0, 4)


def test_catalog_without_lines(wsm, orc):
    c = synth.tiny_case(nl=16, nf=50, np_=4)
    import copy

    cat = copy.deepcopy(c.cat)
    for name in ("f0", "a", "e0", "gu", "gl", "T0"):
        setattr(cat, name, getattr(cat, name)[:0])
    cat.ls_offset = np.zeros(1, np.int64)
    cat.ls_species = cat.ls_species[:0]
    cat.ls_type = cat.ls_type[:0]
    cat.ls_X = cat.ls_X[:0]
    cat.band_offset = np.zeros(len(cat.band_isot) + 1, np.int64)
    for name in ("z_on", "z_gu", "z_gl", "two_Ju", "two_Jl"):
        setattr(cat, name, getattr(cat, name)[:0])
    K, _ = wsm.spectral_propmat_pathFromPath(cat, c.f, c.atm)
    assert not K.any()
    I, _ = wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg)
    assert np.array_equal(I, c.I_bkg)  # K == 0: rotational -> J = 0, T = 1


def test_single_level_path(wsm, orc):
    c = synth.tiny_case(nl=32, nf=100, np_=1)
    I, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    assert np.array_equal(I, c.I_bkg)  # no layer: the background passes through (rtepack_rtestep.cc:287)
    Ir, _ = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    assert np.array_equal(Ir, c.I_bkg)


def test_two_levels_and_one_frequency(wsm, orc):
    c = synth.tiny_case(nl=32, nf=1, np_=2)
    I, _, K = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, return_propmat=True)
    Ir, _, Kr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, return_K=True)
    assert_propmat_close(K, Kr)
    np.testing.assert_allclose(I, Ir, rtol=1e-9)


def test_select_species_without_bands(wsm):
    """A species id with no band selects nothing (lbl_lineshape.cpp:191): K stays zero."""
    c = synth.tiny_case(nl=32, nf=64, np_=2)
    c.cat.n_species = 3  # species 2 exists in the atmosphere but owns no band
    c.atm.vmr = np.concatenate([c.atm.vmr, np.full((c.np_, 1), 1e-3)], axis=1)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, select_species=2)
    assert not K.any()


def test_zero_vmr_and_zero_pressure_levels(wsm, orc):
    """vmr = 0 of the absorber and P = 0 (top of the atmosphere): strengths and widths degenerate
    (y = 0, Doppler only) without NaNs on either side."""
    c = synth.tiny_case(nl=32, nf=200, np_=4)
    c.atm.vmr[0, :] = 0.0
    c.atm.P[1] = 0.0
    Kr, _ = orc.propmat_levels(c.cat, c.f, c.atm)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    assert np.isfinite(K).all() and not K[0].any() and not K[1].any()
    assert_propmat_close(K, Kr)


def test_far_infrared_magnitudes(wsm, orc):
    """Frequencies up to 100 THz: u^2 ~ 1e28 and D^2 ~ 1e56 in the far-wing arithmetic stay finite."""
    c = synth.case_c4(n_lines=300, nf=400, np_=3)
    Kr, _ = orc.propmat_levels(c.cat, c.f, c.atm)
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    assert_propmat_close(K, Kr)


def test_negative_pressure_broadening_is_rejected(wsm):
    c = synth.tiny_case(nl=16, nf=32, np_=2)
    c.cat.ls_X[:, abi.VAR_G0, 0] = -1e4
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    assert e.value.code == abi.ERR_UNSUPPORTED


def test_bad_inputs_are_rejected(wsm):
    c = synth.tiny_case(nl=16, nf=32, np_=2)
    c.atm.T[1] = -5.0
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    assert e.value.code == abi.ERR_INVALID
    c = synth.tiny_case(nl=16, nf=32, np_=2)
    c.atm.Q[0, 0] = 0.0
    with pytest.raises(wsm.Ab200Error):
        wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
