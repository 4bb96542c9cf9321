"""rte_option = linprop for unpolarised layers: tran::linsrc_linprop / linsrc_linprop_deriv
(rtepack_transmission.cc:449-541) with Faddeeva::Dawson, TransmittanceMatrix::linprop (:1195-1252); polarised
layers with a positive absorption gradient need the complex matrix functions of :872-1002 and are rejected."""
import json
import os

import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_linsrc_convergence_fixture_with_linprop(wsm):
    """The reference's own fixture (tests/core/linsrc/test_linsrc_convergence.py) incl. its linprop curve."""
    gold = json.load(open(os.path.join(GOLD, "linsrc_convergence.json")))
    f = np.array([100e9])
    for varying, key in ((False, "constant_k"), (True, "varying")):
        tb = []
        N, scl = 2**12, 1.0
        while N >= 2:
            k = np.linspace(1e-2, 1e-4, N) if varying else np.full(N, 1e-2)
            K = np.zeros((N, 1, 7))
            K[:, 0, 0] = k
            Tlev = np.linspace(200.0, 300.0, N)
            r = np.full(N - 1, scl)
            bkg = np.zeros((1, 4))
            bkg[0, 0] = synth.planck(f, 100.0)[0]
            tm = wsm.spectral_tramat_pathFromPath(K, None, r, Tlev, "linprop")
            J, dJ = wsm.spectral_rad_srcvec_pathFromPropmat(K, f, Tlev)
            I, _ = wsm.spectral_radStepByStepEmission(tm, J, dJ, bkg)
            tb.append(wsm.spectral_radApplyPlanckTb(I, f)[0, 0])
            N //= 2
            scl *= 2
        tb = np.array(tb)
        assert np.abs(tb - np.array(gold[key]["linprop"])).max() <= 1e-6
        lin = np.array(gold[key]["constant"])
        assert np.all(lin / lin[0] >= tb / tb[0])  # :95, :181, :269


def _scalar_K(rng, np_, nf):
    K = np.zeros((np_, nf, 7))
    K[..., 0] = rng.uniform(0.2, 2.0, (np_, nf)) * 1e-4
    K[:, ::4, 0] = np.linspace(0.2e-4, 3e-4, np_)[:, None]  # monotone increasing columns: Dawson branch in every layer
    K[:, 1::4, 0] = np.linspace(3e-4, 0.2e-4, np_)[:, None]  # decreasing: linsrc fallback in every layer
    return K


def test_unfused_linprop_scalar_with_jacobians(wsm, orc):
    rng = np.random.default_rng(12)
    np_, nf, nq = 8, 120, 2
    K = _scalar_K(rng, np_, nf)
    dK = np.zeros((np_, nq, nf, 7))
    dK[..., 0] = rng.uniform(-1, 1, (np_, nq, nf)) * 1e-6
    r = rng.uniform(100.0, 900.0, np_ - 1)
    assert ((K[1:, :, 0] - K[:-1, :, 0]) / (2 * r[:, None]) >= 1e-8).mean() > 0.2  # both branches are exercised
    Tlev = np.linspace(210.0, 290.0, np_)
    f = np.linspace(50e9, 70e9, nf)
    bkg = np.zeros((nf, 4))
    bkg[:, 0] = synth.planck(f, 2.735)
    dr = np.zeros((2, np_ - 1, nq))
    dr[0, :, 0] = r / (2 * Tlev[:-1])
    dr[1, :, 0] = r / (2 * Tlev[1:])
    Tr, Lr, Pr, dTr, dLr = orc.tramat(K, dK, r, dr, "linprop")
    tm = wsm.spectral_tramat_pathFromPath(K, dK, r, Tlev, "linprop", hse_derivative=1, it=0)
    np.testing.assert_allclose(tm.L.reshape(Lr.shape), Lr, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(tm.dL.reshape(dLr.shape), dLr, rtol=1e-7, atol=1e-9 * np.abs(dLr).max())
    # linprop differs from linsrc where the gradient is positive
    _, Ls, _, _, _ = orc.tramat(K, dK, r, dr, "linsrc")
    assert np.abs(Lr - Ls).max() > 1e-6
    Jr, dJr = orc.srcvec(K, f, Tlev, it=0, nq=nq)
    J, dJ = wsm.spectral_rad_srcvec_pathFromPropmat(K, f, Tlev, it=0, nq=nq)
    Ir, dIr = orc.rte_emission("linprop", Tr, Lr, Pr, dTr, dLr, Jr, dJr, bkg)
    I, dI = wsm.spectral_radStepByStepEmission(tm, J, dJ, bkg)
    np.testing.assert_allclose(I, Ir, rtol=1e-10, atol=1e-13 * np.abs(Ir).max())
    np.testing.assert_allclose(dI, dIr, rtol=1e-7, atol=1e-9 * np.abs(dIr).max())


@pytest.mark.parametrize("targets", [(), (("T",), ("VMR", 3))])
def test_fused_linprop_scalar(wsm, orc, targets):
    c = synth.case_c2(lines_per_species=300, nf=900, np_=25, bands_per_species=3, rte_option="linprop")
    c.cat.a[:] *= 100.0  # strong enough that 16 % of the layers have a gradient above the 1e-8 threshold (:456)
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option="linprop", targets=targets, hse_derivative=1)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option="linprop", jac_targets=targets,
                                             hse_derivative=1)
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, c.f), orc.planck_tb(c.f, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6
    Is, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option="linsrc")
    assert np.abs(I - Is).max() > 0, "fixture must reach the Dawson branch"
    for q in range(len(targets)):
        sc = np.abs(dIr[:, :, q, 0]).max()
        assert np.abs(dI[:, :, q, 0] - dIr[:, :, q, 0]).max() <= 5e-7 * sc


def test_polarised_linprop_is_rejected(wsm):
    rng = np.random.default_rng(3)
    np_, nf = 5, 40
    K = np.zeros((np_, nf, 7))
    K[..., 0] = np.linspace(1e-4, 5e-4, np_)[:, None]
    K[..., 1:] = rng.uniform(-1e-5, 1e-5, (np_, nf, 6))
    r = np.full(np_ - 1, 100.0)
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.spectral_tramat_pathFromPath(K, None, r, np.full(np_, 250.0), "linprop")
    assert e.value.code == abi.ERR_UNSUPPORTED
    # decreasing absorption: every layer falls back to the (polarised) linsrc operator -> fine
    tm = wsm.spectral_tramat_pathFromPath(K[::-1].copy(), None, r, np.full(np_, 250.0), "linprop")
    ts = wsm.spectral_tramat_pathFromPath(K[::-1].copy(), None, r, np.full(np_, 250.0), "linsrc")
    assert np.array_equal(tm.L, ts.L)
