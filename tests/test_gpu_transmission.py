"""GPU parity of spectral_radCumulativeTransmission (src/m_spectral_radiance.cc:49-74, rte_transmission
rtepack_rtestep.cc:456-503) against the CPU oracle: un-fused on materialised T, P, dT and fused
(AB200_FLAG_NO_EMISSION on the resident path)."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.test_gpu_jacobian import assert_jac_close

pytestmark = pytest.mark.gpu
TARGETS = (("T",), ("VMR", 0))


def _sun(nf):
    I0 = np.zeros((nf, 4))
    I0[:, 0] = np.linspace(1.0, 2.0, nf) * 1e-12
    I0[:, 1] = 0.1e-12
    I0[:, 3] = -0.05e-12
    return I0


@pytest.mark.parametrize("zeeman", [False, True])
@pytest.mark.parametrize("option", ["constant", "linsrc"])
def test_unfused_transmission(wsm, orc, zeeman, option):
    c = synth.tiny_case(nl=64, nf=38 * 5 if zeeman else 257, np_=6, zeeman=zeeman, rte_option=option)
    K, dK = orc.propmat_levels(c.cat, c.f, c.atm, targets=TARGETS)
    T, L, P, dT, dL = orc.tramat(K, dK, c.r, None, option)
    I0 = _sun(c.nf)
    Ir, dIr = orc.rte_transmission(T, P, dT, I0)
    tm = wsm.TransmittanceMatrix(option, T.reshape(c.nf, c.np_, 4, 4), None, P.reshape(c.nf, c.np_, 4, 4),
                                 dT.reshape(2, c.nf, c.np_, 2, 4, 4), None)
    I, dI = wsm.spectral_radCumulativeTransmission(tm, I0)
    # one matrix-vector product per frequency; the device contracts it into FMAs
    np.testing.assert_allclose(I, Ir, rtol=1e-14, atol=1e-15 * np.abs(Ir).max())
    for q in range(2):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=1e-12, what=f"transmission dI target {q}")
    assert np.abs(dIr).max() > 0
    # no targets: only the forward part
    tm0 = wsm.TransmittanceMatrix(option, tm.T, None, tm.P, None, None)
    I2, dI2 = wsm.spectral_radCumulativeTransmission(tm0, I0)
    assert np.array_equal(I2, I) and dI2.shape == (c.nf, c.np_, 0, 4)
    with pytest.raises(ValueError, match="Bad background radiance size"):
        wsm.spectral_radCumulativeTransmission(tm, I0[:-1])


@pytest.mark.parametrize("zeeman", [False, True])
def test_fused_transmission(wsm, orc, zeeman):
    c = synth.tiny_case(nl=64, nf=38 * 5 if zeeman else 300, np_=7, zeeman=zeeman)
    if not zeeman:
        c.cat.a *= 0.01  # optical depths of 1..6 instead of 70..650: a transmission worth comparing at 1e-9
    I0 = _sun(c.nf)
    K, dK = orc.propmat_levels(c.cat, c.f, c.atm, targets=TARGETS)
    T, L, P, dT, dL = orc.tramat(K, dK, c.r, None, "linsrc")
    Ir, dIr = orc.rte_transmission(T, P, dT, I0)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, I0, jac_targets=TARGETS, flags=abi.FLAG_NO_EMISSION)
    np.testing.assert_allclose(I, Ir, rtol=1e-9, atol=1e-12 * np.abs(Ir).max())
    for q in range(2):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], what=f"fused transmission dI target {q}")
    If, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, I0, flags=abi.FLAG_NO_EMISSION)
    assert np.array_equal(If, I)
    Ie, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, I0)
    assert (Ie[:, 0] > If[:, 0]).all(), "emission adds to the transmitted background"
