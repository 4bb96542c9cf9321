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
