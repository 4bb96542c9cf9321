"""GPU path against the committed oracle-made golden vectors (tests/golden/path_goldens.npz,
generator tests/golden/make_oracle_goldens.py) — the same checks as the live-oracle tests, but
against files that were produced in the container where the reference tree is mounted."""
import os

import numpy as np
import pytest

from tests.conftest import assert_propmat_close
from tests.golden.make_oracle_goldens import golden_cases

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "path_goldens.npz"))


@pytest.mark.parametrize("name", ["c1_small", "c1_cutoff", "c2_small", "c3_small"])
def test_propmat_golden(wsm, name):
    c = golden_cases()[name]
    K, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    assert_propmat_close(K, GOLD[name + "_K"], atol_scale=1e-11 if "cutoff" in name else 1e-12)


@pytest.mark.parametrize("name", ["c2_small", "c3_small"])
@pytest.mark.parametrize("option", ["constant", "linsrc"])
def test_radiance_golden(wsm, name, option):
    c = golden_cases()[name]
    I, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option)
    tb = wsm.spectral_radApplyPlanckTb(I, c.f)
    assert np.abs(tb - GOLD[f"{name}_Tb_{option}"]).max() <= 1e-6  # K, north_star
    np.testing.assert_allclose(I[:, 0], GOLD[f"{name}_I_{option}"][:, 0], rtol=1e-9)


def test_jacobian_golden_every_target_kind(wsm):
    """tests/golden/jacobian_goldens.npz: a small Zeeman path with one target of every kind (temperature, VMR, wind,
    magnetic field, isotopologue ratio, line centre, line-shape coefficient), propagation-matrix rows and radiance rows
    through the fused chain."""
    from tests.golden.make_oracle_goldens import jacobian_case
    from tests.test_gpu_jacobian import assert_jac_close

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "jacobian_goldens.npz"))
    c, tg = jacobian_case()
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    assert_propmat_close(K, g["K"])
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1)
    np.testing.assert_allclose(I, g["I"], rtol=1e-9, atol=1e-12 * np.abs(g["I"]).max())
    for q in range(len(tg)):
        assert np.abs(g["dK"][:, q]).max() > 0 and np.abs(g["dI"][:, :, q]).max() > 0, tg[q]
        assert_jac_close(dK[:, q], g["dK"][:, q], rtol=5e-7, what=f"dK {tg[q]}")
        assert_jac_close(dI[:, :, q], g["dI"][:, :, q], rtol=5e-7, what=f"dI {tg[q]}")
