"""GPU parity of spectral_propmatAddLookup (src/m_lookup.cc:20-173, src/core/lookup/lookup_map.cpp) against the CPU
oracle, and the reference's own kind of check (tests/core/lookup/calc.py): a table precomputed with the line-by-line path on the
GPU reproduces the line-by-line brightness temperatures of perturbed atmospheres."""
import copy

import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.test_oracle_pins import _lut_fixture

pytestmark = pytest.mark.gpu


def _atm_for(tab, rng, n=4):
    P = np.exp(rng.uniform(tab.log_p_grid[-1], tab.log_p_grid[0], n))
    Tref = np.interp(np.log(P), tab.log_p_grid[::-1], tab.t_atmref[::-1])
    wref = np.interp(np.log(P), tab.log_p_grid[::-1], tab.water_atmref[::-1]) if tab.water_atmref is not None else np.full(n, 1e-3)
    return abi.AtmPath(T=Tref + rng.uniform(-10, 10, n), P=P, vmr=np.stack([wref * rng.uniform(0.6, 2.5, n), np.full(n, 0.2)], 1),
                       isorat=np.ones((n, 1)), Q=np.ones((n, 1)))


@pytest.mark.parametrize("do_t,do_w", [(True, True), (True, False), (False, True), (False, False)])
def test_lookup_levels_host_buffers(wsm, orc, do_t, do_w):
    rng = np.random.default_rng(41 + 2 * do_t + do_w)
    tabs = [_lut_fixture(rng, do_t, do_w, species=1), _lut_fixture(rng, do_t, False, species=0, nf=20)]
    tabs[1].f_grid = tabs[0].f_grid[0] + (tabs[1].f_grid - tabs[1].f_grid[0]) * 0.9  # inside the first table's range
    tabs[1].log_p_grid = tabs[0].log_p_grid
    lut = wsm.Lookup(tabs)
    f = np.sort(rng.uniform(max(t.f_grid[0] for t in tabs), min(t.f_grid[-1] for t in tabs), 333))  # inside both tables
    atm = _atm_for(tabs[0], rng)
    tg, d = (("T",), ("VMR", 0), ("VMR", 1)), (0.1, 1e-6, 1e-4)
    for sel in (abi.SPECIES_BATH, 1):
        for orders, noneg in (((3, 2, 3, 1), 1), ((5, 4, 4, 0), 0), ((7, 4, 5, 7), 1)):
            Kr, dKr = orc.lookup_levels(tabs, f, atm, h2o_species=0, select_species=sel, targets=tg, target_d=d, orders=orders,
                                        no_negative_absorption=noneg)
            K = np.zeros((atm.np_, len(f), 7)); dK = np.zeros((atm.np_, 3, len(f), 7))
            wsm.spectral_propmatAddLookup(K, dK, f, tg, sel, lut, atm, h2o_species=0, target_d=d, no_negative_absorption=noneg,
                                          p_interp_order=orders[0], t_interp_order=orders[1], water_interp_order=orders[2],
                                          f_interp_order=orders[3])
            np.testing.assert_allclose(K[..., 0], Kr[..., 0], rtol=1e-10, atol=1e-13 * np.abs(Kr).max())
            sc = np.abs(dKr[..., 0]).max(axis=(0, 2), keepdims=True)
            # the rows are differences of two extractions divided by a small step: 1e-10 of each extraction / d
            assert (np.abs(dK[..., 0] - dKr[..., 0]) <= 1e-6 * np.maximum(sc, 1e-300)).all()
            assert not K[..., 1:].any() and not dK[..., 1:].any()
    far = abi.AtmPath(T=atm.T, P=atm.P * 1e6, vmr=atm.vmr, isorat=np.ones((atm.np_, 1)), Q=np.ones((atm.np_, 1)))
    K = np.zeros((atm.np_, len(f), 7))
    with pytest.raises(wsm.Ab200Error, match="check_limit"):
        wsm.spectral_propmatAddLookup(K, None, f, (), abi.SPECIES_BATH, lut, far, h2o_species=0, p_interp_order=3, t_interp_order=2,
                                      water_interp_order=3, f_interp_order=1)
    if do_t:  # t_pert has five points
        with pytest.raises(wsm.Ab200Error, match="Too few grid points"):
            wsm.spectral_propmatAddLookup(K, None, f, (), abi.SPECIES_BATH, lut, atm, h2o_species=0, p_interp_order=7, t_interp_order=7)
    with pytest.raises(wsm.Ab200Error, match="above 7"):
        wsm.spectral_propmatAddLookup(K, None, f, (), abi.SPECIES_BATH, lut, atm, h2o_species=0, p_interp_order=8)
    lut.close()


def test_precomputed_table_reproduces_line_by_line_radiance(wsm, orc):
    """abs_lookup_dataPrecompute with the GPU line sum (lookup_map.cpp:22-131), then the agenda with use_abs_lookup_data = 1 on
    perturbed atmospheres against the line-by-line run: the reference's tests/core/lookup/calc.py asserts 1e-3 K."""
    c = synth.case_c2(lines_per_species=300, nf=400, np_=41, bands_per_species=2)
    cat = wsm.Catalog(c.cat)
    h2o = 0
    t_pert, w_pert = np.linspace(-30, 30, 7), np.geomspace(0.1, 10, 9)
    ref_atm = c.atm if c.atm.P[0] > c.atm.P[-1] else c.atm.reversed()  # table profiles run surface -> top
    with pytest.raises(ValueError, match="descending pressures"):
        wsm.abs_lookup_dataPrecompute(cat, ref_atm.reversed(), c.f, 1)
    tables = []
    for s in range(c.cat.n_species):
        tables.append(wsm.abs_lookup_dataPrecompute(cat, ref_atm, c.f, s, temperature_perturbation=t_pert,
                                                     water_perturbation=w_pert if s == h2o else None, h2o_species=h2o))
    # the table at zero offset / unit ratio is the line-by-line cross-section itself (against the oracle)
    K0, _ = orc.propmat_levels(c.cat, c.f, ref_atm, select_species=1)
    nd = ref_atm.vmr[:, 1] * ref_atm.P / (1.380649e-23 * ref_atm.T)
    np.testing.assert_allclose(tables[1].xsec[3, 0], K0[..., 0] / nd[:, None], rtol=1e-9)
    # with partition-function tables the builder evaluates Q at the perturbed temperature, like lbl::calculate does
    q_slope = float(ref_atm.Q[0, 0] / ref_atm.T[0])  # the synthetic Q(T) is linear in T
    pf = [("coeff", None, [0.0, q_slope])] * c.cat.n_species
    tq = wsm.abs_lookup_dataPrecompute(cat, ref_atm, c.f, 1, temperature_perturbation=t_pert, partfun_tables=pf)
    cold = copy.deepcopy(ref_atm)
    cold.T = ref_atm.T + t_pert[0]
    cold.Q = ref_atm.Q * (cold.T / ref_atm.T)[:, None]
    Kc, _ = orc.propmat_levels(c.cat, c.f, cold, select_species=1)
    ndc = cold.vmr[:, 1] * cold.P / (1.380649e-23 * cold.T)
    np.testing.assert_allclose(tq.xsec[0, 0], Kc[..., 0] / ndc[:, None], rtol=1e-9)
    assert np.abs(tq.xsec[0, 0] / tables[1].xsec[0, 0] - 1).max() > 0.05, "Q(T) must matter"
    lut = wsm.Lookup(tables)
    worst = 0.0
    for ratio, off in ((0.5, -20.0), (0.5, 10.0), (5.0, 0.0), (5.0, 20.0), (1.0, -10.0)):
        atm = copy.deepcopy(c.atm)
        atm.T = atm.T + off
        atm.vmr = atm.vmr.copy(); atm.vmr[:, h2o] *= ratio
        path = wsm.Path(cat, c.nf, c.np_, 0)
        path.upload(c.f, atm, c.r, c.I_bkg)
        path.run_propmat(); path.run_stokes()
        I = np.empty((c.nf, 4)); path.download(I=I)
        path.add_lookup(lut, h2o_species=h2o, p_interp_order=5, t_interp_order=4, water_interp_order=4, f_interp_order=0)
        path.run_stokes()
        Il = np.empty((c.nf, 4)); Kl = np.empty((c.np_, c.nf, 7)); path.download(I=Il, K=Kl)
        path.close()
        tb, tbl = wsm.spectral_radApplyPlanckTb(I, c.f)[:, 0], wsm.spectral_radApplyPlanckTb(Il, c.f)[:, 0]
        worst = max(worst, float(np.abs(tb - tbl).max()))
        # and the extraction itself against the oracle on the same tables
        Kr, _ = orc.lookup_levels(tables, c.f, atm, h2o_species=h2o, orders=(5, 4, 4, 0))
        np.testing.assert_allclose(Kl[..., 0], Kr[..., 0], rtol=1e-9, atol=1e-12 * Kr.max())
    assert worst < 5e-2, f"LUT vs LBL brightness temperature: {worst} K"
    lut.close(); cat.close()
