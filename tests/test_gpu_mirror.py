"""GPU parity of the VP_LTE_MIRROR line shape (lbl_lineshape_voigt_lte_mirrored.cpp): every sub-line plus
its mirror image at -f0', forward and with the mirrored engine's literal dT / dVMR (:305-325), against the CPU oracle."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.conftest import assert_propmat_close
from tests.test_gpu_jacobian import assert_jac_close

pytestmark = pytest.mark.gpu
TARGETS = (("T",), ("VMR", 0), ("VMR", 1))


def _low_frequency_case(nl=300, nf=700, np_=4, mixed=False):
    """Lines at 1-60 GHz in a thick atmosphere: the mirror images reach a few per cent of the absorption."""
    c = synth.tiny_case(nl=nl, nf=nf, np_=np_)
    rng = np.random.default_rng(21)
    nb = len(c.cat.band_isot)
    for b in range(nb):
        lo, hi = c.cat.band_offset[b], c.cat.band_offset[b + 1]
        c.cat.f0[lo:hi] = np.sort(rng.uniform(1e9, 60e9, hi - lo))
    c.f = np.linspace(0.5e9, 80e9, nf)
    c.cat.band_lineshape[:] = abi.LINESHAPE_VP_LTE_MIRROR
    if mixed:
        c.cat.band_lineshape[::2] = abi.LINESHAPE_VP_LTE  # plain and mirrored bands of one species in one merged segment
    return c


@pytest.mark.parametrize("mixed", [False, True])
def test_mirrored_real_bands(wsm, orc, mixed):
    c = _low_frequency_case(mixed=mixed)
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=TARGETS)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=TARGETS)
    assert_propmat_close(K, Kr)
    for q in range(3):
        for lev in range(c.np_):
            assert_jac_close(dK[lev, q], dKr[lev, q], what=f"mirrored dK target {q} level {lev}")
    plain = synth.tiny_case(nl=300, nf=700, np_=4)
    plain.cat.f0[:] = c.cat.f0
    K0, _ = orc.propmat_levels(plain.cat, c.f, c.atm)
    assert (Kr[..., 0] / K0[..., 0]).max() > 1.01, "the mirror images must contribute"
    cat = wsm.Catalog(c.cat)
    assert cat.counts()[0] == 300, "counts are the reference's sub-lines, not the twin slots"
    cat.close()


def test_mirrored_with_line_mixing_and_zeeman(wsm, orc):
    """Complex segments: line mixing (Y, G) and Zeeman sub-lines, all polarisations, 7 components."""
    c = synth.tiny_case(nf=38 * 8, np_=3, zeeman=True)
    c.cat.band_lineshape[:] = abi.LINESHAPE_VP_LTE_MIRROR
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=TARGETS[:2])
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=TARGETS[:2])
    assert_propmat_close(K, Kr)
    for q in range(2):
        assert_jac_close(dK[:, q], dKr[:, q], rtol=5e-7, what=f"mirrored Zeeman dK target {q}")
    c2 = _low_frequency_case(nl=40, nf=300, np_=3)
    c2.cat.ls_type[:, abi.VAR_Y] = abi.TM_T1
    c2.cat.ls_X[:, abi.VAR_Y, 0] = np.random.default_rng(2).uniform(-3e-6, 3e-6, len(c2.cat.ls_species))
    c2.cat.ls_X[:, abi.VAR_Y, 1] = 0.8
    for clamp in (0, 1):
        Kr, _ = orc.propmat_levels(c2.cat, c2.f, c2.atm, no_negative_absorption=clamp)
        K, _ = wsm.spectral_propmat_pathFromPath(c2.cat, c2.f, c2.atm, no_negative_absorption=clamp)
        assert_propmat_close(K, Kr, atol_scale=1e-11)


def test_mirrored_radiance_and_unsupported_cutoff(wsm, orc):
    c = _low_frequency_case(nl=100, nf=400, np_=6)
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, targets=TARGETS[:2])
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=TARGETS[:2])
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, c.f), orc.planck_tb(c.f, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6
    for q in range(2):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=5e-7, what=f"mirrored dI target {q}")
    c.cat.band_cutoff_type[:] = abi.CUTOFF_BYLINE
    c.cat.band_cutoff_value[:] = 5e9
    with pytest.raises(wsm.Ab200Error, match="VP_LTE_MIRROR with a cutoff"):
        wsm.Catalog(c.cat)
