"""GPU parity of the line-parameter Jacobians (compute_derivative(line_key),
lbl_lineshape_voigt_lte.cpp:1562-1637): f0, e0, Einstein coefficient and the line-shape model coefficients of ONE
catalog line, against the CPU oracle (itself pinned to perturbed catalogs, tests/test_oracle_pins.py)."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.conftest import assert_propmat_close
from tests.test_gpu_jacobian import assert_jac_close
from tests.test_oracle_pins import _line_target_fixture

pytestmark = pytest.mark.gpu


def _all_targets(line, sp0):
    tg = [("line_f0", line), ("line_e0", line), ("line_a", line)]
    for var in (abi.VAR_G0, abi.VAR_D0, abi.VAR_DV, abi.VAR_Y, abi.VAR_G):
        for k in (0, 1):
            tg += [("line_ls", line, var, sp0, k), ("line_ls", line, var, abi.SPECIES_BATH, k)]
    return tg


def test_line_targets_real_bands(wsm, orc):
    c, line = _line_target_fixture()
    c.f = np.linspace(c.cat.f0.min() - 2e9, c.cat.f0.max() + 2e9, 900)  # near and far tiles, other bands in between
    tg = _all_targets(line, int(c.cat.ls_species[c.cat.ls_offset[line]]))
    for i in range(0, len(tg), 8):
        part = tg[i:i + 8] + ([("T",)] if i == 0 else [])[: 8 - len(tg[i:i + 8])]
        Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=part)
        K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=part)
        assert_propmat_close(K, Kr)
        for q in range(len(part)):
            assert np.abs(dKr[:, q]).max() > 0, part[q]
            assert_jac_close(dK[:, q], dKr[:, q], what=f"dK target {part[q]}")


@pytest.mark.parametrize("tm", [abi.TM_T0, abi.TM_T2, abi.TM_T3, abi.TM_T4, abi.TM_T5, abi.TM_AER, abi.TM_DPL, abi.TM_POLY])
def test_line_shape_coefficients_of_every_temperature_model(wsm, orc, tm):
    c, line = _line_target_fixture()
    lo = c.cat.ls_offset[line]
    rng = np.random.default_rng(int(tm))
    c.cat.ls_type[lo:lo + 2, abi.VAR_G0] = tm
    x = c.cat.ls_X[lo:lo + 2, abi.VAR_G0]
    if tm == abi.TM_POLY:
        x[:, 1:] = x[:, :1] * np.array([1e-3, -1e-6, 1e-9])
    elif tm == abi.TM_AER:
        x[:, 1:] = x[:, :1] * rng.uniform(0.8, 1.2, (2, 3))
    elif tm == abi.TM_DPL:
        x[:, 1:] = np.array([0.7, 1.0, 0.4]) * np.array([[1.0, x[0, 0] * 0.1, 1.0], [1.0, x[1, 0] * 0.1, 1.0]])
    elif tm == abi.TM_T3:
        x[:, 1] = x[:, 0] * 1e-3
    else:
        x[:, 1:] = rng.uniform(0.2, 0.9, (2, 3))
    tg = [("line_ls", line, abi.VAR_G0, sp, k) for sp in (int(c.cat.ls_species[lo]), abi.SPECIES_BATH) for k in range(4)]
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    assert_propmat_close(K, Kr)
    for q in range(8):
        assert_jac_close(dK[:, q], dKr[:, q], what=f"model {tm} target {tg[q]}")


def test_line_targets_zeeman_line_through_the_fused_chain(wsm, orc):
    c = synth.case_c3(nf=38 * 8, np_=4, los=(140.0, 25.0))
    line = 7
    tg = [("line_f0", line), ("line_ls", line, abi.VAR_Y, abi.SPECIES_BATH, 0), ("T",), ("line_ls", line, abi.VAR_G0, abi.SPECIES_BATH, 1),
          ("line_a", line)]
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    for q in range(len(tg)):
        assert np.abs(dKr[:, q]).max() > 0, tg[q]
        assert_jac_close(dK[:, q], dKr[:, q], rtol=5e-7, what=f"Zeeman dK target {tg[q]}")
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, targets=tg, hse_derivative=1)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1)
    for q in range(len(tg)):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=5e-7, what=f"dI target {tg[q]}")


def test_line_target_errors(wsm, orc):
    c = synth.tiny_case(nl=16, nf=32, np_=2, cutoff=1e9)
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=[("line_f0", 3)])
    assert e.value.code == abi.ERR_UNSUPPORTED  # the reference indexes the cutoff window with whole-band indices
    with pytest.raises(RuntimeError):
        orc.propmat_levels(c.cat, c.f, c.atm, targets=[("line_f0", 3)])
    c = synth.tiny_case(nl=16, nf=32, np_=2)
    with pytest.raises(wsm.Ab200Error) as e:  # AtmKey::p, lbl_lineshape_voigt_lte.cpp:1482
        wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=[("p",)])
    assert e.value.code == abi.ERR_UNSUPPORTED and "pressure derivative" in str(e.value)
    with pytest.raises(RuntimeError, match="pressure derivative"):
        orc.propmat_levels(c.cat, c.f, c.atm, targets=[("p",)])
    for bad in (("line_a", 16), ("line_a", -1), ("line_ls", 2, 9, 0, 0), ("line_ls", 2, abi.VAR_G0, 0, 4), ("line_ls", 2, abi.VAR_G0, 77, 0)):
        with pytest.raises(wsm.Ab200Error) as e:
            wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=[bad])
        assert e.value.code == abi.ERR_INVALID, bad
