"""GPU parity of the magnetic-field Jacobians: d propmat / d mag_{u,v,w} of the Zeeman polarisations
(lbl_lineshape_voigt_lte.cpp:1066-1162, :1484-1513: splitting derivative s dz dF through norm_view plus dnorm_view times
the absorption itself), forward through the Stokes chain like any other target, against the CPU oracle."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.conftest import assert_propmat_close
from tests.test_gpu_jacobian import assert_jac_close

pytestmark = pytest.mark.gpu
MAG = (("mag_u",), ("mag_v",), ("mag_w",))


def test_magnetic_jacobian_propmat_zeeman(wsm, orc):
    c = synth.case_c3(nf=38 * 8, np_=4, los=(140.0, 25.0))
    tg = (("mag_u",), ("T",), ("mag_w",), ("VMR", 0), ("mag_v",))  # two passes of the Jacobian kernel
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    assert_propmat_close(K, Kr)
    for q in (0, 2, 4):
        assert (np.abs(dKr[:, q]).reshape(-1, 7).max(axis=0) > 0).all(), "all seven components respond to the field"
    for q in range(len(tg)):
        assert_jac_close(dK[:, q], dKr[:, q], rtol=5e-7, what=f"dK target {tg[q]}")


def test_magnetic_jacobian_with_cutoff_and_unsplit_lines(wsm, orc):
    """A Zeeman band with a ByLine cutoff next to plain bands: the pol = no pass adds nothing to a magnetic row."""
    c = synth.tiny_case(nf=38 * 8, np_=3, zeeman=True, cutoff=2e9)
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=MAG)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=MAG)
    assert_propmat_close(K, Kr, atol_scale=1e-11)
    for q in range(3):
        assert_jac_close(dK[:, q], dKr[:, q], rtol=5e-7, what=f"dK magnetic target {q}")
    plain = synth.tiny_case(nl=64, nf=200, np_=3)  # no Zeeman line at all, field present: rows stay zero
    plain.atm.mag = np.tile(np.array([1e-5, -2e-5, 3e-5]), (3, 1))
    plain.atm.los = np.tile(np.array([120.0, 10.0]), (3, 1))
    _, dK0 = wsm.spectral_propmat_pathFromPath(plain.cat, plain.f, plain.atm, jac_targets=MAG)
    assert not dK0.any()


@pytest.mark.parametrize("option", ["constant", "linsrc"])
def test_magnetic_jacobian_through_the_fused_chain(wsm, orc, option):
    c = synth.case_c3(nf=38 * 8, np_=6, los=(150.0, 40.0), rte_option=option)
    tg = (("mag_v",), ("T",), ("mag_w",), ("mag_u",))
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, targets=tg, hse_derivative=1, rte_option=option)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1,
                                             rte_option=option)
    np.testing.assert_allclose(I, Ir, rtol=1e-9, atol=1e-12 * np.abs(Ir).max())
    assert np.abs(dIr[:, :, 0]).max() > 0
    for q in range(4):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=5e-7, what=f"dI target {tg[q]}")


def test_magnetic_jacobian_mirrored_zeeman_band(wsm, orc):
    c = synth.tiny_case(nf=38 * 8, np_=3, zeeman=True)
    c.cat.band_lineshape[:] = abi.LINESHAPE_VP_LTE_MIRROR
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=MAG[:2])
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=MAG[:2])
    assert_propmat_close(K, Kr)
    for q in range(2):
        assert_jac_close(dK[:, q], dKr[:, q], rtol=5e-7, what=f"mirrored dK magnetic target {q}")
