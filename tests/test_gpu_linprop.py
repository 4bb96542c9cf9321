"""rte_option = linprop: tran::linsrc_linprop / linsrc_linprop_deriv (rtepack_transmission.cc:449-556)
with Faddeeva::Dawson, TransmittanceMatrix::linprop (:1195-1252).  Unpolarised layers take the scalar Dawson form and
its closed derivative; polarised layers with an absorption gradient take the complex matrix square root (:872-1002), its
inverse and the element-wise complex Dawson function, with a 1e-6 perturbation for the derivative (:543-555)."""
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


def _polarised_K(rng, np_, nf, pol=0.2):
    """absorption rising along the path (positive gradient in most layers) with a polarised part of relative size pol"""
    a = 10 ** rng.uniform(-5.5, -3.7, (1, nf, 1)) * (1.0 + np.arange(np_)[:, None, None] * rng.uniform(0.3, 2.0, (1, nf, 1)))
    K = a * np.concatenate([np.ones((np_, nf, 1)), rng.uniform(-pol, pol, (np_, nf, 6))], axis=2)
    return np.ascontiguousarray(K)


def test_unfused_polarised_linprop_with_jacobians(wsm, orc):
    rng = np.random.default_rng(31)
    np_, nf, nq = 7, 150, 2
    K = _polarised_K(rng, np_, nf)
    K[:, ::7, 1:] = 0.0    # unpolarised columns in between
    K[3:, 5::11, 0] *= 0.2  # and layers with a falling absorption: the linsrc fall-back (:456)
    dK = K[:, None] * rng.uniform(-1e-2, 1e-2, (np_, nq, nf, 7))
    r = 10 ** rng.uniform(1.0, 2.5, np_ - 1)
    grad = (K[1:, :, 0] - K[:-1, :, 0]) / (2 * r[:, None]) >= 1e-8
    assert 0.3 < grad.mean() < 0.99
    Tlev = np.linspace(210.0, 290.0, np_)
    f = np.linspace(50e9, 70e9, nf)
    dr = np.zeros((2, np_ - 1, nq))
    dr[0, :, 0] = r / (2 * Tlev[:-1])
    dr[1, :, 0] = r / (2 * Tlev[1:])
    Tr, Lr, Pr, dTr, dLr = orc.tramat(K, dK, r, dr, "linprop")
    tm = wsm.spectral_tramat_pathFromPath(K, dK, r, Tlev, "linprop", hse_derivative=1, it=0)
    L = tm.L.reshape(Lr.shape)
    sc = np.abs(Lr).max(axis=2, keepdims=True)  # per (frequency, level): the matrix is ~ Lambda * identity + small terms
    assert np.abs(L - Lr).max() <= 1e-11 * sc.max()
    np.testing.assert_allclose(L, Lr, rtol=0, atol=1e-11 * float(sc.max()))
    _, Ls, _, _, _ = orc.tramat(K, dK, r, dr, "linsrc")
    assert np.abs(Lr - Ls).max() > 1e-6  # the Dawson form is not the linsrc operator
    # the polarised derivative is a forward difference with eps = 1e-6 of Lambda itself: errors of Lambda (1e-13) show at 1e-7
    dL = tm.dL.reshape(dLr.shape)
    assert np.abs(dL - dLr).max() <= 3e-6 * np.abs(dLr).max() + 1e-6 * 1e-11 / 1e-6
    bkg = np.zeros((nf, 4))
    bkg[:, 0] = synth.planck(f, 2.735)
    Jr, dJr = orc.srcvec(K, f, Tlev, it=0, nq=nq)
    J, dJ = wsm.spectral_rad_srcvec_pathFromPropmat(K, f, Tlev, it=0, nq=nq)
    Ir, dIr = orc.rte_emission("linprop", Tr, Lr, Pr, dTr, dLr, Jr, dJr, bkg)
    I, dI = wsm.spectral_radStepByStepEmission(tm, J, dJ, bkg)
    np.testing.assert_allclose(I, Ir, rtol=0, atol=1e-11 * np.abs(Ir).max())
    assert np.abs(dI - dIr).max() <= 3e-6 * np.abs(dIr).max()
    # decreasing absorption everywhere: every layer falls back to the (polarised) linsrc operator
    Kd = np.ascontiguousarray(_polarised_K(rng, np_, 40)[::-1])
    tl = wsm.spectral_tramat_pathFromPath(Kd, None, r, Tlev, "linprop")
    ts = wsm.spectral_tramat_pathFromPath(Kd, None, r, Tlev, "linsrc")
    assert np.array_equal(tl.L, ts.L)


@pytest.mark.parametrize("targets", [(), (("T",), ("mag_u",))])
def test_fused_polarised_linprop_zeeman(wsm, orc, targets):
    """Zeeman-split O2 lines (full 4x4 propagation matrix), looking down so that the absorption rises along the path towards
    the surface: the fused chain (first column of Lambda only: 8 complex Dawson evaluations per layer) and the fused
    Jacobian pass against the oracle's un-fused restatement."""
    c = synth.case_c3(nf=38 * 12, np_=9, los=(140.0, 30.0))
    a = c.atm

    def rev(x):
        return None if x is None else np.ascontiguousarray(x[::-1])

    # sensor above the path (level 0 = 80 km): the absorption rises away from the sensor; thin layers keep the gradient
    # (K_{i+1} - K_i) / 2r above the reference's 1e-8 threshold in 45 % of the (frequency, layer) pairs
    atm = abi.AtmPath(T=rev(a.T), P=rev(a.P), vmr=rev(a.vmr), isorat=rev(a.isorat), Q=rev(a.Q), dQdT=rev(a.dQdT), mag=rev(a.mag),
                      los=rev(a.los))
    r = np.ascontiguousarray(c.r[::-1]) * 0.02
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, atm, r, c.I_bkg, rte_option="linprop", targets=targets, hse_derivative=0)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, atm, r, c.I_bkg, rte_option="linprop", jac_targets=targets)
    Is, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, atm, r, c.I_bkg, rte_option="linsrc")
    assert np.abs(I - Is).max() > 1e-2 * np.abs(I).max(), "fixture must reach the polarised Dawson branch"
    np.testing.assert_allclose(I, Ir, rtol=0, atol=1e-10 * np.abs(Ir).max())
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, c.f), orc.planck_tb(c.f, Ir)
    assert np.abs(tb[:, 0] - tbr[:, 0]).max() <= 1e-6
    for q in range(len(targets)):
        sc = np.abs(dIr[:, :, q, :]).max()
        assert np.abs(dI[:, :, q, :] - dIr[:, :, q, :]).max() <= 1e-5 * sc
