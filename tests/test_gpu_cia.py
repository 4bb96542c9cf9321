"""GPU parity of spectral_propmatAddCIA (src/m_cia.cc:27-178, src/core/absorption/cia.cc:76-226) against the
CPU oracle: host-buffer form, on the resident path (lines + CIA -> radiance), Jacobian rows, error behaviour."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.test_oracle_pins import _cia_fixture

pytestmark = pytest.mark.gpu
TARGETS = (("T",), ("VMR", 0), ("VMR", 1))


def _atm(n=4):
    T = np.linspace(210.0, 330.0, n)
    return abi.AtmPath(T=T, P=np.geomspace(9e4, 4e3, n), vmr=np.tile([0.78, 0.21, 4e-4], (n, 1)), isorat=np.ones((n, 1)),
                       Q=np.ones((n, 1)))


def test_cia_levels_host_buffers(wsm, orc):
    recs = _cia_fixture(np.random.default_rng(12))
    f = np.linspace(0.5e11, 2.2e12, 1777)
    atm = _atm()
    cia = wsm.Cia(recs)
    for sel in (abi.SPECIES_BATH, 0, 2, 1):
        Kr, dKr = orc.cia_levels(recs, f, atm, select_species=sel, targets=TARGETS, dT=0.1)
        K = np.zeros((atm.np_, len(f), 7))
        dK = np.zeros((atm.np_, 3, len(f), 7))
        wsm.spectral_propmatAddCIA(K, dK, f, TARGETS, sel, cia, atm, dT=0.1)
        np.testing.assert_allclose(K[..., 0], Kr[..., 0], rtol=1e-11, atol=1e-13 * max(Kr.max(), 1e-300))
        sc = np.abs(dKr[..., 0]).max(axis=(0, 2), keepdims=True)
        assert (np.abs(dK[..., 0] - dKr[..., 0]) <= 1e-9 * np.maximum(sc, 1e-300)).all()
        assert not K[..., 1:].any() and not dK[..., 1:].any(), "only A is touched"
        K2 = np.full((atm.np_, len(f), 7), 0.5)  # += semantics on the caller's arrays (m_cia.cc:146-148)
        wsm.spectral_propmatAddCIA(K2, None, f, (), sel, cia, atm)
        np.testing.assert_allclose(K2[..., 0], 0.5 + Kr[..., 0], rtol=1e-15)
        assert np.array_equal(K2[..., 1:], np.full_like(K2[..., 1:], 0.5))
        if sel == 1:
            assert not Kr.any(), "select_species filters on the FIRST species of the pair (m_cia.cc:105-106)"
    # per-level frequency grids
    f2 = f[None, :] * (1 + 1e-5 * np.arange(atm.np_)[:, None])
    Kr, _ = orc.cia_levels(recs, f2, atm)
    K = np.zeros((atm.np_, len(f), 7))
    wsm.spectral_propmatAddCIA(K, None, f2, (), abi.SPECIES_BATH, cia, atm)
    np.testing.assert_allclose(K[..., 0], Kr[..., 0], rtol=1e-11, atol=1e-13 * Kr.max())
    # a wind target would need the re-extraction at f + df (m_cia.cc:78-81): refused, not ignored
    with pytest.raises(wsm.Ab200Error, match="wind target") as e:
        wsm.spectral_propmatAddCIA(np.zeros((atm.np_, len(f), 7)), np.zeros((atm.np_, 1, len(f), 7)), f, (("wind_u",),), abi.SPECIES_BATH, cia, atm)
    assert e.value.code == abi.ERR_UNSUPPORTED
    cia.close()


def test_cia_temperature_extrapolation_error_and_nan(wsm, orc):
    recs = _cia_fixture(np.random.default_rng(12))
    cia = wsm.Cia(recs)
    f = np.linspace(0.5e11, 2.2e12, 300)
    cold = abi.AtmPath(T=[250.0, 120.0], P=[5e4, 1e4], vmr=np.tile([0.78, 0.21, 4e-4], (2, 1)), isorat=np.ones((2, 1)), Q=np.ones((2, 1)))
    K = np.zeros((2, len(f), 7))
    with pytest.raises(wsm.Ab200Error, match="extrapolation range"):
        wsm.spectral_propmatAddCIA(K, None, f, (), abi.SPECIES_BATH, cia, cold)
    Kr, _ = orc.cia_levels(recs, f, cold, ignore_errors=1)
    wsm.spectral_propmatAddCIA(K, None, f, (), abi.SPECIES_BATH, cia, cold, ignore_errors=1)
    assert np.array_equal(np.isnan(K[..., 0]), np.isnan(Kr[..., 0])) and np.isnan(Kr[1, :, 0]).all()
    np.testing.assert_allclose(K[0, :, 0], Kr[0, :, 0], rtol=1e-11, atol=1e-13 * Kr[0].max())
    # frequencies that miss the failing data set entirely: no error (cia.cc:95-118 returns before the temperature is looked at)
    f_lo = np.linspace(0.5e11, 1.9e11, 50)
    cold1 = abi.AtmPath(T=[120.0], P=[1e4], vmr=[[0.78, 0.21, 4e-4]], isorat=[[1.0]], Q=[[1.0]])
    K1 = np.zeros((1, 50, 7))
    wsm.spectral_propmatAddCIA(K1, None, f_lo, (), 2, cia, cold1)  # pair (2, 2) has a single temperature: no limit at all
    Kr1, _ = orc.cia_levels(recs, f_lo, cold1, select_species=2)
    np.testing.assert_allclose(K1[..., 0], Kr1[..., 0], rtol=1e-11)
    with pytest.raises(wsm.Ab200Error, match="Not enough frequency grid points"):
        wsm.Cia([abi.CiaRecord(0, 0, [(np.array([1.0, 2, 3]), np.array([200.0]), np.ones((3, 1)))])])
    cia.close()


def test_cia_on_the_resident_path(wsm, orc):
    """Lines + CIA in the same K, then the fused Stokes chain and its Jacobians: the agenda's order, m_abs.cc:257-296."""
    c = synth.tiny_case(nl=64, nf=400, np_=6, targets=TARGETS[:2])
    rng = np.random.default_rng(3)
    fg = np.linspace(90e9, 140e9, 30)
    Tg = np.array([180.0, 230.0, 280.0, 330.0])
    dat = 3e-50 * (1 + 0.2 * rng.normal(size=(30, 4))) * np.exp(-((fg - 115e9) / 15e9) ** 2)[:, None]
    recs = [abi.CiaRecord(0, 1, [(fg, Tg, dat)])]
    tg = TARGETS[:2]
    # oracle: un-fused chain with CIA added to K / dK
    K, dK = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K0 = K.copy()
    orc.cia_levels(recs, c.f, c.atm, targets=tg, dT=0.1, K=K, dK=dK)
    assert (K[..., 0] - K0[..., 0]).max() > 0.05 * K0[..., 0].max(), "the CIA term must matter in this fixture"
    T, L, P, dT, dL = orc.tramat(K, dK, c.r, None, "linsrc")
    J, dJ = orc.srcvec(K, c.f, c.atm.T, 0, 2)
    Ir, dIr = orc.rte_emission("linsrc", T, L, P, dT, dL, J, dJ, c.I_bkg)
    cat, cia = wsm.Catalog(c.cat), wsm.Cia(recs)
    path = wsm.Path(cat, c.nf, c.np_, 2)
    path.upload(c.f, c.atm, c.r, c.I_bkg, targets=tg)
    path.run_propmat()
    path.add_cia(cia, dT=0.1)
    path.run_stokes()
    I = np.empty((c.nf, 4)); dI = np.empty((c.nf, c.np_, 2, 4)); Kg = np.empty((c.np_, c.nf, 7))
    path.download(I=I, dI=dI, K=Kg)
    np.testing.assert_allclose(Kg[..., 0], K[..., 0], rtol=1e-9)
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, c.f), orc.planck_tb(c.f, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6
    from tests.test_gpu_jacobian import assert_jac_close
    for q in range(2):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=5e-7, what=f"lines + CIA dI target {q}")
    path.close(); cia.close(); cat.close()
