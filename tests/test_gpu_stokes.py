"""GPU parity of stage 2 (rtepack Stokes chain), fused and un-fused, against the CPU oracle
(reference src/core/rtepack/rtepack_transmission.cc, rtepack_source.cc, rtepack_rtestep.cc)."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth

pytestmark = pytest.mark.gpu
TB_TOL = 1e-6  # K, north_star


def _random_K(rng, np_, nf, polarised, scale=1e-4):
    K = np.zeros((np_, nf, 7))
    K[..., 0] = rng.uniform(0.2, 2.0, (np_, nf)) * scale
    if polarised:
        K[..., 1:] = rng.uniform(-0.3, 0.3, (np_, nf, 6)) * scale
        K[:, ::5, 1:] = 0.0  # unpolarised columns mixed in
        K[:, 3::7, 4:] = 0.0  # y == 0 branch
        K[:, 4::7, 1:4] = 0.0  # purely rotational off-diagonal
        K[1, 10:20, :4] = 0.0  # is_rotational() -> J = 0
    return K


@pytest.mark.parametrize("polarised", [False, True])
@pytest.mark.parametrize("option", ["constant", "linsrc"])
@pytest.mark.parametrize("flags", [0, abi.FLAG_TRAN_EXACT])
def test_unfused_chain_matches_oracle(wsm, orc, polarised, option, flags):
    rng = np.random.default_rng(3)
    np_, nf = 9, 301
    K = _random_K(rng, np_, nf, polarised)
    r = rng.uniform(200.0, 3000.0, np_ - 1)
    f = np.linspace(50e9, 70e9, nf)
    Tlev = np.linspace(210.0, 290.0, np_)
    bkg = np.zeros((nf, 4))
    bkg[:, 0] = synth.planck(f, 2.735)
    bkg[:, 1] = 0.01 * bkg[:, 0]

    Tr, Lr, Pr, _, _ = orc.tramat(K, None, r, None, option, flags)
    tm = wsm.spectral_tramat_pathFromPath(K, None, r, Tlev, option, flags=flags)
    for got, ref, name in ((tm.T, Tr, "T"), (tm.P, Pr, "P")) + (((tm.L, Lr, "L"),) if option == "linsrc" else ()):
        np.testing.assert_allclose(got.reshape(ref.shape), ref, rtol=1e-11, atol=1e-13, err_msg=name)
    assert np.array_equal(tm.T[:, 0].reshape(nf, 16), np.tile(np.eye(4).ravel(), (nf, 1)))

    Jr, _ = orc.srcvec(K, f, Tlev)
    J, dJ = wsm.spectral_rad_srcvec_pathFromPropmat(K, f, Tlev)
    np.testing.assert_allclose(J, Jr, rtol=1e-14)

    Ir, _ = orc.rte_emission(option, Tr, Lr, Pr, np.zeros((2, nf, np_, 0, 16)), np.zeros((2, nf, np_, 0, 16)), Jr,
                             np.zeros((nf, np_, 0, 4)), bkg)
    I, _ = wsm.spectral_radStepByStepEmission(tm, J, dJ, bkg)
    np.testing.assert_allclose(I, Ir, rtol=1e-11, atol=1e-13 * np.abs(Ir).max())


@pytest.mark.parametrize("option", ["constant", "linsrc"])
def test_fused_scalar_clearsky_tb(wsm, orc, option):
    """Reduced BASELINE config 2: 5 species, scalar K, 30 levels; Tb within 1e-6 K."""
    c = synth.case_c2(lines_per_species=400, nf=3000, np_=30, rte_option=option, bands_per_species=4)
    Ir, _, Kr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option, return_K=True)
    I, _, K = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option, return_propmat=True)
    from tests.conftest import assert_propmat_close

    assert_propmat_close(K, Kr)
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, c.f), orc.planck_tb(c.f, Ir)
    assert np.abs(tb - tbr).max() <= TB_TOL
    assert tbr[:, 0].max() - tbr[:, 0].min() > 5.0, "fixture must have spectral structure"
    np.testing.assert_allclose(I, Ir, rtol=1e-9)


@pytest.mark.parametrize("option", ["constant", "linsrc"])
@pytest.mark.parametrize("flags", [0, abi.FLAG_TRAN_EXACT])
def test_fused_zeeman_full_stokes_tb(wsm, orc, option, flags):
    """Reduced BASELINE config 3: polarised propmat, non-diagonal exp(-K r), 4-Stokes Tb within 1e-6 K."""
    c = synth.case_c3(nf=38 * 30, np_=20, los=(120.0, 30.0), rte_option=option)
    Ir, _ = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option, flags=flags)
    I, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option, flags=flags)
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, c.f), orc.planck_tb(c.f, Ir)
    assert np.abs(tb - tbr).max() <= TB_TOL
    assert np.abs(tbr[:, 1:]).max() > 1e-3, "fixture must be polarised"


def test_linsrc_convergence_fixture(wsm, orc):
    """tests/core/linsrc/test_linsrc_convergence.py of the reference (catalog free): constant k and
    linearly varying k, T = linspace(200, 300, N), f = 100 GHz, surface 100 K; orderings :95-96,:181-182."""
    f = np.array([100e9])
    for varying in (False, True):
        lin, linsrc = [], []
        N, scl = 2**12, 1.0
        while N >= 2:
            k = np.linspace(1e-2, 1e-4, N) if varying else np.full(N, 1e-2)
            K = np.zeros((N, 1, 7))
            K[:, 0, 0] = k
            Tlev = np.linspace(200.0, 300.0, N)
            r = np.full(N - 1, scl)
            bkg = np.zeros((1, 4))
            bkg[0, 0] = synth.planck(f, 100.0)[0]
            out = {}
            for opt in ("linsrc", "constant"):
                tm = wsm.spectral_tramat_pathFromPath(K, None, r, Tlev, opt)
                J, dJ = wsm.spectral_rad_srcvec_pathFromPropmat(K, f, Tlev)
                I, _ = wsm.spectral_radStepByStepEmission(tm, J, dJ, bkg)
                out[opt] = wsm.spectral_radApplyPlanckTb(I, f)[0, 0]
                # the oracle on the same inputs
                Tr, Lr, Pr, dTr, dLr = orc.tramat(K, None, r, None, opt)
                Jr, dJr = orc.srcvec(K, f, Tlev)
                Ir, _ = orc.rte_emission(opt, Tr, Lr, Pr, dTr, dLr, Jr, dJr, bkg)
                assert abs(out[opt] - orc.planck_tb(f, Ir)[0, 0]) <= TB_TOL
            linsrc.append(out["linsrc"])
            lin.append(out["constant"])
            N //= 2
            scl *= 2
        lin, linsrc = np.array(lin), np.array(linsrc)
        assert np.all(lin / lin[0] >= linsrc / linsrc[0])


def test_shard_invariance_bitwise(wsm):
    """Frequency partitioning must not change a single bit (north_star): any contiguous shard of the
    grid gives the same spectral_rad as the full run."""
    c = synth.case_c2(lines_per_species=300, nf=4000, np_=12, bands_per_species=3)
    cat = wsm.Catalog(c.cat)
    I, _ = wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg)
    for lo, hi in ((0, 1000), (1000, 1777), (1777, 4000), (513, 514)):
        Is, _ = wsm.spectral_radClearskyEmission(cat, c.f[lo:hi], c.atm, c.r, c.I_bkg[lo:hi])
        assert np.array_equal(Is, I[lo:hi]), (lo, hi)
    c3 = synth.case_c3(nf=38 * 20, np_=6)
    cat3 = wsm.Catalog(c3.cat)
    I3, _ = wsm.spectral_radClearskyEmission(cat3, c3.f, c3.atm, c3.r, c3.I_bkg)
    for lo, hi in ((0, 300), (300, 760)):
        Is, _ = wsm.spectral_radClearskyEmission(cat3, c3.f[lo:hi], c3.atm, c3.r, c3.I_bkg[lo:hi])
        assert np.array_equal(Is, I3[lo:hi]), (lo, hi)
    cat.close()
    cat3.close()


def test_transparent_and_opaque_limits(wsm):
    """Size-independent properties: K = 0 leaves the background untouched (J = 0 for rotational K);
    an opaque isothermal path returns the Planck function of its temperature."""
    c = synth.tiny_case(nl=32, nf=200, np_=8)
    c.cat.a[:] = 0.0
    I, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    assert np.array_equal(I, c.I_bkg)
    c = synth.tiny_case(nl=32, nf=200, np_=8)
    c.cat.a[:] *= 1e9
    c.atm.T[:] = 255.0
    c.atm.Q[:] = 215.0 * 255.0 / 296.0
    I, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    tb = wsm.spectral_radApplyPlanckTb(I, c.f)
    assert np.abs(tb[:, 0] - 255.0).max() < 1e-9
