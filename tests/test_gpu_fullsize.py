"""Full BASELINE sizes on the GPU, checked through size-independent properties (the oracle cannot run
1e12..1e14 evaluations): the per-frequency result does not depend on the grid it is computed in, so

* a strided sample of the full-size spectrum must equal, BIT FOR BIT, a separate GPU run on just the
  sampled frequencies (frequency-partition invariance at full size), and
* that small run is compared with the CPU oracle on the same inputs (<= 1e-9 relative on K.A, <= 1e-6 K on Tb).
"""
import numpy as np
import pytest

from arts_b200 import synth
from tests.conftest import assert_propmat_close

pytestmark = pytest.mark.gpu


def _check(wsm, orc, c, n_sample, option="linsrc"):
    cat = wsm.Catalog(c.cat)
    path = wsm.Path(cat, c.nf, c.np_)
    path.upload(c.f, c.atm, c.r, c.I_bkg, rte_option=option)
    path.run_propmat()
    path.run_stokes()
    I = np.empty((c.nf, 4))
    path.download(I=I)
    idx = np.unique(np.linspace(0, c.nf - 1, n_sample).astype(np.int64))
    fs, bs = np.ascontiguousarray(c.f[idx]), np.ascontiguousarray(c.I_bkg[idx])
    Is, _, Ks = wsm.spectral_radClearskyEmission(cat, fs, c.atm, c.r, bs, rte_option=option, return_propmat=True)
    assert np.array_equal(Is, I[idx]), "full-size run differs from the run on the sampled grid"
    Ir, _, Kr = orc.clearsky_emission(c.cat, fs, c.atm, c.r, bs, rte_option=option, return_K=True)
    assert_propmat_close(Ks, Kr)
    tb, tbr = wsm.spectral_radApplyPlanckTb(Is, fs), orc.planck_tb(fs, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6
    assert np.isfinite(I).all() and (I[:, 0] > 0).all()
    path.close()
    cat.close()
    return tbr


def test_config2_full_size(wsm, orc):
    """BASELINE configs[1]: 5 species x 2e4 lines, 1e5 frequencies, 100 levels (1e12 evaluations)."""
    tb = _check(wsm, orc, synth.case_c2(), n_sample=48)
    assert tb[:, 0].max() - tb[:, 0].min() > 5.0


def test_config3_full_size(wsm, orc):
    """BASELINE configs[2]: O2 60 GHz Zeeman, 50 levels x 1e5 frequencies, polarised chain."""
    c = synth.case_c3()
    assert c.nf == 100_000 and c.np_ == 50
    tb = _check(wsm, orc, c, n_sample=400)
    assert np.abs(tb[:, 1:]).max() > 1e-3


def test_config4_frequency_shard(wsm, orc):
    """BASELINE configs[3] catalog (1e6 lines, 100 levels) on a 65 536-point contiguous shard of the
    1e6-point grid — the per-GPU shape of a 16-way split; 6.5e12 evaluations."""
    c = synth.case_c4(f_slice=(500_000, 565_536))
    assert c.cat.n_lines == 1_000_000 and c.nf == 65_536 and c.np_ == 100
    _check(wsm, orc, c, n_sample=24)


def test_repeated_steps_are_bit_identical(wsm):
    """Run-to-run determinism at the bench shape (1e7 Stokes steps per pass): the TMA ring of the chain kernel once
    refilled a stage before every lane's shared-memory load had returned (a few wrong lanes per pass, found with
    tools/det_check.py); repeated passes must give the same bits."""
    c = synth.case_c2()
    cat = wsm.Catalog(c.cat)
    path = wsm.Path(cat, c.nf, c.np_)
    ref = None
    for option in ("linsrc", "constant"):
        path.upload(c.f, c.atm, c.r, c.I_bkg, rte_option=option)
        path.run_propmat()
        outs = []
        for _ in range(4):
            path.run_stokes()
            I = np.empty((c.nf, 4))
            path.download(I=I)
            outs.append(I)
        for I in outs[1:]:
            assert np.array_equal(I, outs[0]), option
    c3 = synth.case_c3()
    cat3 = wsm.Catalog(c3.cat)
    p3 = wsm.Path(cat3, c3.nf, c3.np_)
    p3.upload(c3.f, c3.atm, c3.r, c3.I_bkg)
    p3.run_propmat()
    outs = []
    for _ in range(3):
        p3.run_stokes()
        I = np.empty((c3.nf, 4))
        p3.download(I=I)
        outs.append(I)
    assert np.array_equal(outs[1], outs[0]) and np.array_equal(outs[2], outs[0])
    for p in (path, p3):
        p.close()
    cat.close()
    cat3.close()


def test_config5_path_with_jacobians_full_size(wsm, orc):
    """BASELINE configs[4], one path of the batch at its full size: 1e4 lines x 1e4 frequencies x 100 levels with
    temperature and VMR Jacobian rows (hse_derivative on), fused chain.  The spectrum and the rows of a strided frequency
    sample must equal a separate GPU run on just those frequencies bit for bit (the very-far / near split of the Jacobian
    line sum is decided per pair), and that run is compared with the oracle: K <= 1e-9, Tb <= 1e-6 K, rows <= 2e-7 of the
    column maximum."""
    c = synth.case_c5_single()
    assert c.cat.n_lines == 10_000 and c.nf == 10_000 and c.np_ == 100
    tg = (("T",), ("VMR", 0))
    cat = wsm.Catalog(c.cat)
    I, dI = wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1)
    idx = np.unique(np.linspace(0, c.nf - 1, 40).astype(np.int64))
    fs, bs = np.ascontiguousarray(c.f[idx]), np.ascontiguousarray(c.I_bkg[idx])
    Is, dIs, Ks = wsm.spectral_radClearskyEmission(cat, fs, c.atm, c.r, bs, jac_targets=tg, hse_derivative=1, return_propmat=True)
    assert np.array_equal(Is, I[idx]) and np.array_equal(dIs, dI[idx])
    Ir, dIr, Kr = orc.clearsky_emission(c.cat, fs, c.atm, c.r, bs, targets=tg, hse_derivative=1, return_K=True)
    assert_propmat_close(Ks, Kr)
    tb, tbr = wsm.spectral_radApplyPlanckTb(Is, fs), orc.planck_tb(fs, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6
    for q in range(len(tg)):
        a, b = dIs[:, :, q, 0], dIr[:, :, q, 0]
        assert np.abs(b).max() > 0
        assert np.abs(a - b).max() <= 2e-7 * np.abs(b).max(), (q, np.abs(a - b).max() / np.abs(b).max())
    cat.close()
