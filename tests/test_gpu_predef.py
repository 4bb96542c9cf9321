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
        wsm.spectral_propmatAddPredefined(K2, None, [99], abi.SPECIES_BATH, (), f, atm, SPECIES)
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


FULL_MODELS = ["H2O-PWR98", "O2-PWR98", "H2O-MPM89", "O2-MPM89", "N2-SelfContMPM93"]
PWR20XX = ["H2O-PWR2021", "O2-PWR2021", "N2-SelfContPWR2021", "H2O-PWR2022", "O2-PWR2022", "O2-TRE05", "O2-MPM2020"]


@pytest.mark.parametrize("models", [["H2O-PWR98", "O2-PWR98", "N2-SelfContMPM93"], ["H2O-MPM89", "O2-MPM89"], FULL_MODELS + MODELS,
                                    ["H2O-PWR2021", "O2-PWR2021", "N2-SelfContPWR2021"], ["H2O-PWR2022", "O2-PWR2022"], PWR20XX + MODELS])
def test_full_microwave_models_host_buffers(wsm, orc, models):
    """PWR98 (H2O, O2), MPM89 (H2O, O2), MPM93 N2, PWR2021 / PWR2022 (H2O with its speed-dependent cores through the device's
    w(i z), O2 with second-order mixing, N2): the per-level line tables the CTA builds in shared memory against the
    oracle's per-frequency restatement (itself pinned bit for bit to the reference's object code, tests/test_refslice_pins.py),
    every level of a 9-level profile, 1-1000 GHz with the 60 GHz band resolved, temperature + three VMR rows."""
    f = np.sort(np.concatenate([np.linspace(1e9, 1e12, 1500), np.linspace(50e9, 70e9, 700), np.linspace(118.2e9, 119.3e9, 60)]))
    atm = _atm(9)
    tg, d = (("T",), ("VMR", 0), ("VMR", 1), ("VMR", 2)), (0.1, 1e-6, 1e-4, 1e-4)
    for sel in (abi.SPECIES_BATH, 0, 1):
        Kr, dKr = orc.predef_levels(models, SPECIES, f, atm, select_species=sel, targets=tg, target_d=d)
        K = np.zeros((atm.np_, len(f), 7)); dK = np.zeros((atm.np_, 4, len(f), 7))
        wsm.spectral_propmatAddPredefined(K, dK, models, sel, tg, f, atm, SPECIES, target_d=d)
        sc = np.abs(Kr[..., 0]).max(axis=1, keepdims=True)
        # 1e-9 of the element where the absorption is not a near-cancelling sum of positive and negative mixing terms
        assert (np.abs(K[..., 0] - Kr[..., 0]) <= 1e-11 * np.abs(Kr[..., 0]) + 1e-14 * sc).all()
        for q in range(4):
            err, dsc = np.abs(dK[:, q, :, 0] - dKr[:, q, :, 0]).max(), np.abs(dKr[:, q, :, 0]).max()
            assert err <= 2e-6 * dsc, (models, sel, q, err, dsc)
        assert not K[..., 1:].any() and not dK[..., 1:].any()
    assert Kr[..., 0].max() > 0


def test_full_o2_models_refuse_a_tiny_o2_mixing_ratio(wsm):
    f = np.linspace(50e9, 70e9, 300)
    atm = _atm(3)
    atm.vmr[1, 1] = 1e-26
    K = np.zeros((3, len(f), 7))
    for m in ("O2-PWR98", "O2-MPM89", "O2-TRE05"):
        with pytest.raises(wsm.Ab200Error, match="below the threshold"):
            wsm.spectral_propmatAddPredefined(K, None, [m], abi.SPECIES_BATH, (), f, atm, SPECIES)
    atm.vmr[1, 1] = 0.0  # exactly zero: the models add nothing at that level (PWR98.cc:359-361)
    wsm.spectral_propmatAddPredefined(K, None, ["O2-PWR98", "O2-MPM89"], abi.SPECIES_BATH, (), f, atm, SPECIES)
    assert not K[1].any() and K[0, :, 0].min() > 0


def test_microwave_window_tb_with_full_models(wsm, orc):
    """A clear-sky microwave spectrum made of the full models alone (no catalog lines): K from PWR98 + MPM93 on the resident
    path, fused Stokes chain, brightness temperatures within 1e-6 K of the oracle's un-fused chain."""
    c = synth.tiny_case(nl=8, nf=900, np_=12, targets=())
    f = np.linspace(10e9, 200e9, 900)
    species = {"H2O": 0, "O2": 1}
    models = ["H2O-PWR98", "O2-PWR98"]
    K = np.zeros((c.np_, 900, 7))
    orc.predef_levels(models, species, f, c.atm, K=K)
    T, L, P, dT, dL = orc.tramat(K, None, c.r, None, "linsrc")
    J, dJ = orc.srcvec(K, f, c.atm.T, -1, 0)
    bkg = np.zeros((900, 4)); bkg[:, 0] = synth.planck(f, 2.725)
    Ir, _ = orc.rte_emission("linsrc", T, L, P, dT, dL, J, dJ, bkg)
    cat = wsm.Catalog(c.cat)
    path = wsm.Path(cat, 900, c.np_, 0)
    path.upload(f, c.atm, c.r, bkg)
    path.run_propmat()
    path.add_predefined(models, species)
    path.run_stokes()
    I = np.empty((900, 4)); Kg = np.empty((c.np_, 900, 7))
    path.download(I=I, K=Kg)
    Kl, _ = orc.propmat_levels(c.cat, f, c.atm)
    np.testing.assert_allclose(Kg[..., 0], K[..., 0] + Kl[..., 0], rtol=1e-9)
    K2 = K + Kl
    T, L, P, dT, dL = orc.tramat(K2, None, c.r, None, "linsrc")
    J, dJ = orc.srcvec(K2, f, c.atm.T, -1, 0)
    Ir, _ = orc.rte_emission("linsrc", T, L, P, dT, dL, J, dJ, bkg)
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, f), orc.planck_tb(f, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6
    assert tb[:, 0].max() - tb[:, 0].min() > 20.0  # the 22, 60, 118 and 183 GHz features are there
    path.close(); cat.close()


def test_wind_rows_of_the_continua(wsm, orc):
    """freq_jac of PredefinedModel::compute (predefined_absorption_models.cc:280-296): a wind target's row is the frequency
    derivative (model(f + d) - model(f)) / d.  Alone (spectral_propmatAddPredefined) it stays d/df; on the resident path it gets
    the wind fix of the line rows, times f times freq_wind_shift_jac (spectral_propmat_jacWindFix), unless the caller asked for
    d/df rows."""
    f = np.linspace(20e9, 200e9, 901)
    atm = _atm(5)
    models = ["H2O-PWR98", "O2-PWR98", "O2-SelfContStandardType", "H2O-PWR2022"]
    tg, d = (("wind_v",), ("T",), ("wind_u",)), (1e3, 0.1, 2e3)
    Kr, dKr = orc.predef_levels(models, SPECIES, f, atm, targets=tg, target_d=d)
    K = np.zeros((atm.np_, len(f), 7)); dK = np.zeros((atm.np_, 3, len(f), 7))
    wsm.spectral_propmatAddPredefined(K, dK, models, abi.SPECIES_BATH, tg, f, atm, SPECIES, target_d=d)
    np.testing.assert_allclose(K[..., 0], Kr[..., 0], rtol=1e-11)
    for q in range(3):
        sc = np.abs(dKr[:, q, :, 0]).max()
        assert sc > 0 and np.abs(dK[:, q, :, 0] - dKr[:, q, :, 0]).max() <= 1e-5 * sc, q  # a 1e3 Hz difference quotient of 1e-15 noise
    # independent check of the oracle row: centred difference of the model in frequency
    Kp, _ = orc.predef_levels(models, SPECIES, f + 5e5, atm)
    Km, _ = orc.predef_levels(models, SPECIES, f - 5e5, atm)
    cd = (Kp[..., 0] - Km[..., 0]) / 1e6
    assert np.abs(dKr[:, 0, :, 0] - cd).max() <= 2e-3 * np.abs(cd).max()  # narrow O2 lines at 20 hPa against a 1 MHz stencil

    # resident path: lines + continua with a wind target, against lines alone + the oracle's d/df row times f * freq_wind_shift_jac
    c = synth.tiny_case(nl=64, nf=400, np_=6, targets=(("wind_u",),))
    c.atm.wind = np.tile([12.0, -7.0, 1.0], (c.np_, 1))
    c.atm.los = np.tile([130.0, 25.0], (c.np_, 1))
    species = {"H2O": 0, "O2": 1}
    mdl = ["H2O-PWR98", "O2-PWR98"]
    cat = wsm.Catalog(c.cat)

    def run(with_continua, flags=0):
        path = wsm.Path(cat, c.nf, c.np_, 1)
        path.upload(c.f, c.atm, c.r, c.I_bkg, targets=(("wind_u",),), flags=flags)
        path.run_propmat()
        if with_continua:
            path.add_predefined(mdl, species, target_d=(1e3,))
        Kg = np.empty((c.np_, c.nf, 7)); dKg = np.empty((c.np_, 1, c.nf, 7))
        path.download(K=Kg, dK=dKg)
        path.close()
        return Kg, dKg

    (K0, dK0), (K1, dK1) = run(False), run(True)
    fac_jac = [orc.wind_shift(c.atm.wind[i], c.atm.los[i]) for i in range(c.np_)]
    f_lev = np.stack([fj[0] * c.f for fj in fac_jac])  # the shifted grid of every level
    _, dref = orc.predef_levels(mdl, species, f_lev, c.atm, targets=(("wind_u",),), target_d=(1e3,))
    want = dref[:, 0, :, 0] * f_lev * np.array([fj[1][0] for fj in fac_jac])[:, None]
    got = dK1[:, 0, :, 0] - dK0[:, 0, :, 0]
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    (_, dK0f), (_, dK1f) = run(False, abi.FLAG_WIND_ROWS_DF), run(True, abi.FLAG_WIND_ROWS_DF)
    assert np.abs((dK1f - dK0f)[:, 0, :, 0] - dref[:, 0, :, 0]).max() <= 1e-5 * np.abs(dref[:, 0, :, 0]).max()
    cat.close()


def test_liquid_cloud_ell07(wsm, orc):
    """"liquidcloud-ELL07" (src/core/predefined/ELL07.cc:39-188): liquid water absorption from Ellison's 2007 permittivity; the
    species' "mixing ratio" is the liquid water content [kg/m3].  Forward and the T / liquidcloud / wind rows against the oracle
    (which is pinned bit for bit to the reference's ELL07.cc slice), composed with the gas models, selection by species, levels
    without cloud, and the reference's range errors."""
    sp = {"H2O": 0, "O2": 1, "N2": 2, "liquidcloud": 3}
    n = 6
    atm = _atm(n)
    lwc = np.array([0.0, 2e-4, 4.9e-3, 3e-5, 9.9e-11, 1e-10])  # (the perturbed points must stay inside the model's range too)
    atm.vmr = np.ascontiguousarray(np.concatenate([atm.vmr, lwc[:, None]], 1))
    f = np.concatenate([np.linspace(1e9, 1e12, 700), np.geomspace(1e12, 24.9e12, 200)])
    tg, d = (("T",), ("VMR", 3), ("wind_u",), ("VMR", 0)), (0.1, 1e-7, 1e3, 1e-6)
    for models, sel in ((["liquidcloud-ELL07"], abi.SPECIES_BATH), (["liquidcloud-ELL07", "H2O-PWR98", "O2-PWR98"], abi.SPECIES_BATH),
                        (["liquidcloud-ELL07", "H2O-PWR98"], 3), (["liquidcloud-ELL07", "H2O-PWR98"], 0)):
        ff = f if len(models) == 1 else f[:700]
        Kr, dKr = orc.predef_levels(models, sp, ff, atm, select_species=sel, targets=tg, target_d=d)
        K = np.zeros((n, len(ff), 7)); dK = np.zeros((n, 4, len(ff), 7))
        wsm.spectral_propmatAddPredefined(K, dK, models, sel, tg, ff, atm, sp, target_d=d)
        sc = np.abs(Kr[..., 0]).max(axis=1, keepdims=True)
        assert (np.abs(K[..., 0] - Kr[..., 0]) <= 1e-12 * np.abs(Kr[..., 0]) + 1e-14 * sc).all()
        for q in range(4):
            dsc = np.abs(dKr[:, q, :, 0]).max()
            # rows are difference quotients over d: the 1e-14 agreement of the model values themselves is divided by d
            tol = 1e-5 * dsc + 1e-13 * np.abs(Kr[..., 0]).max() / abs(d[q])
            assert np.abs(dK[:, q, :, 0] - dKr[:, q, :, 0]).max() <= max(tol, 1e-300), (models, sel, q)
        assert not K[..., 1:].any() and not dK[..., 1:].any()
        if models == ["liquidcloud-ELL07"]:
            assert not K[0].any() and not K[4].any() and K[5, :, 0].max() > 0  # nothing below 1e-10 kg/m3
            assert K[2, :, 0].max() > K[1, :, 0].max() > K[3, :, 0].max() > 0
            # the liquidcloud row of a linear model is the model per unit content; level 4 crosses the 1e-10 threshold with the step
            np.testing.assert_allclose(dK[1, 1, :, 0] * lwc[1], K[1, :, 0], rtol=1e-6)
            assert dK[4, 1, :, 0].max() > 0 and not dK[0, 0].any()
    K = np.zeros((n, len(f), 7))
    for T0, w0, ff, tgb, db in ((295.0, 5.001e-3, f, (), ()), (209.0, 1e-4, f, (), ()), (374.0, 1e-4, f, (), ()),
                                (295.0, 1e-4, np.append(f, 25.1e12), (), ()), (295.0, 5e-3, f, (("VMR", 3),), (1e-7,)),
                                (372.95, 1e-4, f, (("T",),), (0.1,)), (295.0, 1e-4, np.append(f, 25e12), (("wind_w",),), (1e3,))):
        bad = _atm(n)
        bad.vmr = np.ascontiguousarray(np.concatenate([bad.vmr, lwc[:, None]], 1))
        bad.T[1], bad.vmr[1, 3] = T0, w0
        Kb = np.zeros((n, len(ff), 7)); dKb = np.zeros((n, len(tgb), len(ff), 7))
        with pytest.raises(wsm.Ab200Error, match="ELL07"):  # at the point itself, or at a perturbed point of a Jacobian row
            wsm.spectral_propmatAddPredefined(Kb, dKb if tgb else None, ["liquidcloud-ELL07"], abi.SPECIES_BATH, tgb, ff, bad, sp, target_d=db)
        with pytest.raises(Exception, match="ELL07"):
            orc.predef_levels(["liquidcloud-ELL07"], sp, ff, bad, targets=tgb, target_d=db)
    cold = _atm(n)
    cold.vmr = np.ascontiguousarray(np.concatenate([cold.vmr, np.zeros((n, 1))], 1))
    cold.T[:] = 150.0  # no liquid water anywhere: the range checks are never reached
    wsm.spectral_propmatAddPredefined(K, None, ["liquidcloud-ELL07"], abi.SPECIES_BATH, (), f, cold, sp)
    assert not K.any()
    with pytest.raises(wsm.Ab200Error, match="does not carry"):
        wsm.spectral_propmatAddPredefined(K, None, ["liquidcloud-ELL07"], abi.SPECIES_BATH, (), f, _atm(n), SPECIES)


def test_mt_ckd_water_continua(wsm, orc):
    """"H2O-ForeignContCKDMT400" / "H2O-SelfContCKDMT400" and the 4.3 pair (src/core/predefined/MT_CKD400.cc:102-256, MT_CKD430.cc):
    table-driven water continua.  The reference walks the coefficient table with a cursor along the grid; the device finds every
    frequency's interval on its own - checked against the oracle's sequential restatement (pinned bit for bit to the reference's
    object code) on grids that start below / inside / near the end of the table, per-level grids, with T / H2O / wind rows, composed
    with other models on the resident path, and without data ("No data")."""
    w = synth.mtckd_table()
    w2 = synth.mtckd_table(seed=6, n=1500, v0=-20.0, dv=12.5)
    data = wsm.PredefData(ckdmt400=w, ckdmt430=w2)
    kay = 100 * 299792458.0
    atm = _atm(5)
    tags = ["H2O-ForeignContCKDMT400", "H2O-SelfContCKDMT400", "H2O-ForeignContCKDMT430", "H2O-SelfContCKDMT430"]
    tg, d = (("T",), ("VMR", 0), ("wind_u",), ("VMR", 1)), (0.1, 1e-6, 1e4, 1e-4)
    rng = np.random.default_rng(3)
    wn = w["wavenumbers"]
    grids = [np.linspace(1e9, 6.2e14, 4000), np.sort(rng.uniform(2e12, 1.2e14, 1500)), np.linspace(wn[-4] * kay, wn[-1] * kay * 1.0005, 300),
             np.concatenate([[0.0, wn[2] * kay, wn[3] * kay], np.linspace(1e12, 2e12, 50)]),
             np.stack([np.linspace(3e12, 9e13, 700) * (1 + 1e-5 * i) for i in range(5)])]  # the last one: a grid per level
    for f in grids:
        for models in (tags, tags[1:2]):
            Kr, dKr = orc.predef_levels(models, SPECIES, f, atm, targets=tg, target_d=d, ckdmt400=w, ckdmt430=w2)
            nf = f.shape[-1]
            K = np.zeros((atm.np_, nf, 7)); dK = np.zeros((atm.np_, 4, nf, 7))
            wsm.spectral_propmatAddPredefined(K, dK, models, abi.SPECIES_BATH, tg, f, atm, SPECIES, target_d=d, data=data)
            sc = np.abs(Kr[..., 0]).max(axis=1, keepdims=True)
            assert (np.abs(K[..., 0] - Kr[..., 0]) <= 1e-12 * np.abs(Kr[..., 0]) + 1e-15 * sc).all()
            for q in range(4):
                dsc = np.abs(dKr[:, q, :, 0]).max()
                tol = 1e-5 * dsc + 1e-13 * np.abs(Kr[..., 0]).max() / abs(d[q])
                assert np.abs(dK[:, q, :, 0] - dKr[:, q, :, 0]).max() <= max(tol, 1e-300), (models, q)
            assert not dK[:, 3].any() and not K[..., 1:].any()  # O2 is not a variable of these models: exact zeros
        assert Kr[..., 0].max() > 0
    # on the resident path, after the lines, with a gas model next to them; species selection keeps or drops them
    c = synth.tiny_case(nl=64, nf=400, np_=6, targets=(("T",), ("VMR", 0)))
    species = {"H2O": 0, "O2": 1}
    models = ["H2O-SelfContCKDMT400", "O2-PWR98", "H2O-ForeignContCKDMT430"]
    for sel in (abi.SPECIES_BATH, 0, 1):
        K, dK = orc.propmat_levels(c.cat, c.f, c.atm, targets=c.targets, select_species=sel)
        orc.predef_levels(models, species, c.f, c.atm, select_species=sel, targets=c.targets, target_d=(0.1, 1e-6), K=K, dK=dK, ckdmt400=w, ckdmt430=w2)
        cat = wsm.Catalog(c.cat)
        path = wsm.Path(cat, c.nf, c.np_, 2)
        path.upload(c.f, c.atm, c.r, c.I_bkg, targets=c.targets, select_species=sel)
        path.run_propmat()
        path.add_predefined(models, species, target_d=(0.1, 1e-6), data=data)
        Kg = np.empty_like(K); dKg = np.empty_like(dK)
        path.download(K=Kg, dK=dKg)
        path.close()
        cat.close()
        np.testing.assert_allclose(Kg[..., 0], K[..., 0], rtol=1e-9)
        for q in range(2):
            assert np.abs(dKg[:, q, :, 0] - dK[:, q, :, 0]).max() <= 1e-5 * np.abs(dK[:, q, :, 0]).max()
    K = np.zeros((atm.np_, 10, 7))
    with pytest.raises(wsm.Ab200Error, match="No data"):
        wsm.spectral_propmatAddPredefined(K, None, tags[:1], abi.SPECIES_BATH, (), np.linspace(1e12, 2e12, 10), atm, SPECIES)
    only400 = wsm.PredefData(ckdmt400=w)
    with pytest.raises(wsm.Ab200Error, match="No data"):
        wsm.spectral_propmatAddPredefined(K, None, tags[2:3], abi.SPECIES_BATH, (), np.linspace(1e12, 2e12, 10), atm, SPECIES, data=only400)
    only400.close()
    bad = dict(w, wavenumbers=w["wavenumbers"][::-1].copy())
    with pytest.raises(wsm.Ab200Error, match="increasing"):
        wsm.PredefData(ckdmt400=bad)
    with pytest.raises(wsm.Ab200Error, match="shorter than 4"):
        wsm.PredefData(ckdmt400={k: (v[:3] if hasattr(v, "__len__") else v) for k, v in w.items()})
    data.close()
