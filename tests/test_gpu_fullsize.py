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
