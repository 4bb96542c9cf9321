"""GPU parity of spectral_propmatAddPredefined for the four "StandardType" continua (SURVEY.md 8(f)-2;
src/m_predefined_absorption_models.cc:156-191, src/core/predefined/standard.cc) against the CPU oracle."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth

pytestmark = pytest.mark.gpu
MODELS = ["O2-SelfContStandardType", "N2-SelfContStandardType", "H2O-ForeignContStandardType", "H2O-SelfContStandardType"]
SPECIES = {"H2O": 0, "O2": 1, "N2": 2}


def _atm(n=5):
    return abi.AtmPath(T=np.linspace(295.0, 215.0, n), P=np.geomspace(1e5, 2e3, n),
                       vmr=np.stack([np.geomspace(1.5e-2, 5e-6, n), np.full(n, 0.2095), np.full(n, 0.7808)], 1),
                       isorat=np.ones((n, 1)), Q=np.ones((n, 1)))


def test_predef_levels_host_buffers(wsm, orc):
    f = np.linspace(1e9, 1e12, 1234)
    atm = _atm()
    tg, d = (("T",), ("VMR", 0), ("VMR", 1), ("VMR", 2)), (0.1, 1e-6, 1e-4, 1e-4)
    for sel in (abi.SPECIES_BATH, 0, 1, 2):
        Kr, dKr = orc.predef_levels(MODELS, SPECIES, f, atm, select_species=sel, targets=tg, target_d=d)
        K = np.zeros((atm.np_, len(f), 7)); dK = np.zeros((atm.np_, 4, len(f), 7))
        wsm.spectral_propmatAddPredefined(K, dK, MODELS, sel, tg, f, atm, SPECIES, target_d=d)
        np.testing.assert_allclose(K[..., 0], Kr[..., 0], rtol=1e-12)
        sc = np.abs(dKr[..., 0]).max(axis=(0, 2), keepdims=True)
        # rows are (model(x + d) - model(x)) / d: 1e-15 of the model value over a step of 1e-6
        assert (np.abs(dK[..., 0] - dKr[..., 0]) <= 1e-6 * np.maximum(sc, 1e-300)).all()
        assert not K[..., 1:].any() and not dK[..., 1:].any()
    K2 = np.full((atm.np_, len(f), 7), 0.25)
    wsm.spectral_propmatAddPredefined(K2, None, MODELS, abi.SPECIES_BATH, (), f, atm, SPECIES)
    Kr, _ = orc.predef_levels(MODELS, SPECIES, f, atm)
    np.testing.assert_allclose(K2[..., 0], 0.25 + Kr[..., 0], rtol=1e-15)
    with pytest.raises(wsm.Ab200Error, match="outside the GPU path"):
        wsm.spectral_propmatAddPredefined(K2, None, [17], abi.SPECIES_BATH, (), f, atm, SPECIES)
    with pytest.raises(wsm.Ab200Error, match="does not carry"):
        wsm.spectral_propmatAddPredefined(K2, None, MODELS, abi.SPECIES_BATH, (), f, atm, {"O2": 1, "N2": 2})


def test_lines_plus_continua_on_the_resident_path(wsm, orc):
    """Lines + continua in the same K, then the fused Stokes chain with Jacobians (the agenda order of src/m_abs.cc:257-296)."""
    c = synth.tiny_case(nl=64, nf=400, np_=6, targets=(("T",), ("VMR", 0)))
    tg, d = (("T",), ("VMR", 0)), (0.1, 1e-6)
    species = {"H2O": 0, "O2": 1}
    models = ["O2-SelfContStandardType", "H2O-ForeignContStandardType", "H2O-SelfContStandardType"]
    K, dK = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K0 = K.copy()
    orc.predef_levels(models, species, c.f, c.atm, targets=tg, target_d=d, K=K, dK=dK)
    assert (K[..., 0] - K0[..., 0]).max() > 0
    T, L, P, dT, dL = orc.tramat(K, dK, c.r, None, "linsrc")
    J, dJ = orc.srcvec(K, c.f, c.atm.T, 0, 2)
    Ir, dIr = orc.rte_emission("linsrc", T, L, P, dT, dL, J, dJ, c.I_bkg)
    cat = wsm.Catalog(c.cat)
    path = wsm.Path(cat, c.nf, c.np_, 2)
    path.upload(c.f, c.atm, c.r, c.I_bkg, targets=tg)
    path.run_propmat()
    path.add_predefined(models, species, target_d=d)
    path.run_stokes()
    I = np.empty((c.nf, 4)); dI = np.empty((c.nf, c.np_, 2, 4)); Kg = np.empty((c.np_, c.nf, 7))
    path.download(I=I, dI=dI, K=Kg)
    np.testing.assert_allclose(Kg[..., 0], K[..., 0], rtol=1e-9)
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, c.f), orc.planck_tb(c.f, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6
    from tests.test_gpu_jacobian import assert_jac_close
    for q in range(2):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=5e-7, what=f"lines + continua dI target {q}")
    path.close(); cat.close()
