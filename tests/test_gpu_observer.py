"""GPU parity of the observer epilogue (the callers' glue around the path) against the CPU oracle:
background from a temperature (src/m_background.cc:55-141), spectral_rad_jacFromBackground and
spectral_rad_jacAddPathPropagation (src/m_rad.cc:26-127) accumulated inside the fused Jacobian pass,
spectral_rad_transform_operator (spectral_radiance_transform_operator.cc:8-122) and SensorObsel::sumup
(src/core/sensor/obsel.cpp:246-279).  The oracle side runs the reference's un-fused sequence
propmat -> tramat -> srcvec -> rte_emission -> orc_observer on the host."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.test_gpu_jacobian import assert_jac_close

pytestmark = pytest.mark.gpu

TARGETS = (("T",), ("VMR", 0))


def _observer(c, nq, unit, rng, bkg_T=288.0, n_grid=4, n_channels=5, surface_rows=True):
    """1-D retrieval grid of n_grid nodes per target, linear flat weights at every path point, then one
    surface-temperature row; channels with 20 random frequencies each."""
    nx = n_grid * nq + 1
    path_map = []
    for ip in range(c.np_):
        pos = ip * (n_grid - 1) / max(c.np_ - 1, 1)
        i0 = min(int(pos), n_grid - 2)
        w1 = pos - i0
        path_map.append([[(t * n_grid + i0, 1.0 - w1), (t * n_grid + i0 + 1, w1)] for t in range(nq)])  # w1 == 0 at ip == 0
    channels = []
    for _ in range(n_channels):
        js = np.sort(rng.choice(c.nf, min(20, c.nf), replace=False))
        channels.append([(int(j), (float(rng.uniform(0, 1)), *map(float, rng.uniform(-0.3, 0.3, 3)))) for j in js])
    return abi.Observer(nx=nx, path_map=path_map, bkg_T=bkg_T, bkg_rows=[(nx - 1, 1.0)] if surface_rows and bkg_T else [],
                        unit=unit, n_real=1.00027, channels=channels)


def _oracle(orc, c, targets, obs):
    nq = len(targets)
    K, dK = orc.propmat_levels(c.cat, c.f, c.atm, targets=targets)
    T, L, P, dT, dL = orc.tramat(K, dK, c.r, None, c.rte_option)
    it = [i for i, t in enumerate(targets) if t[0] == "T"]
    J, dJ = orc.srcvec(K, c.f, c.atm.T, it[0] if it else -1, nq)
    I_bkg = c.I_bkg if obs.bkg_T is None else orc.background(c.f, obs.bkg_T)[0]
    I, dI = orc.rte_emission(c.rte_option, T, L, P, dT, dL, J, dJ, I_bkg)
    return orc.observer(c.f, obs, P, I, dI if nq else None)


def _gpu(wsm, c, targets, obs):
    cat = wsm.Catalog(c.cat)
    path = wsm.Path(cat, c.nf, c.np_, len(targets))
    path.upload(c.f, c.atm, c.r, c.I_bkg if obs.bkg_T is None else None, rte_option=c.rte_option, targets=targets)
    path.run_propmat()
    path.run_observer(obs)
    out = path.download_observer()
    path.close()
    cat.close()
    return out


def _compare(got, ref, jac_rtol=2e-7):
    I, Jx, y, Jy = got
    Ir, Jxr, yr, Jyr = ref
    np.testing.assert_allclose(I, Ir, rtol=1e-9, atol=1e-12 * np.abs(Ir).max())
    np.testing.assert_allclose(y, yr, rtol=1e-9, atol=1e-12 * np.abs(yr).max())
    if Jx is not None:
        assert_jac_close(Jx, Jxr, rtol=jac_rtol, what="spectral_rad_jac")
        scale = np.abs(Jyr).max()
        assert np.abs(Jy - Jyr).max() <= jac_rtol * scale, np.abs(Jy - Jyr).max() / scale


@pytest.mark.parametrize("unit", ["unit", "RJBT", "PlanckBT", "W_m2_m_sr", "W_m2_m1_sr"])
def test_observer_scalar_all_units(wsm, orc, unit):
    c = synth.tiny_case(nl=64, nf=257, np_=6, targets=TARGETS)
    obs = _observer(c, 2, unit, np.random.default_rng(1))
    ref = _oracle(orc, c, TARGETS, obs)
    got = _gpu(wsm, c, TARGETS, obs)
    _compare(got, ref)
    assert np.abs(ref[1][-1]).max() > 0, "the surface-temperature row must be exercised"
    assert np.abs(ref[3]).max() > 0


@pytest.mark.parametrize("option", ["constant", "linsrc"])
def test_observer_polarised_background_transmittance(wsm, orc, option):
    """Zeeman-split path: P[np-1] is a full Mueller matrix, so the background row has Q, U, V parts."""
    c = synth.tiny_case(nf=38 * 6, np_=5, zeeman=True, rte_option=option, targets=TARGETS)
    obs = _observer(c, 2, "PlanckBT", np.random.default_rng(2), bkg_T=250.0)
    ref = _oracle(orc, c, TARGETS, obs)
    got = _gpu(wsm, c, TARGETS, obs)
    _compare(got, ref, jac_rtol=5e-7)
    assert np.abs(ref[1][-1][:, 1:]).max() > 0


def test_observer_surface_row_only_and_uploaded_background(wsm, orc):
    c = synth.tiny_case(nl=40, nf=130, np_=4)
    rng = np.random.default_rng(4)
    # no atmospheric target: the Jacobian pass still runs for the cumulative transmittance of the surface row
    obs = _observer(c, 0, "RJBT", rng)
    _compare(_gpu(wsm, c, (), obs), _oracle(orc, c, (), obs))
    # the caller's own background radiance: no background Jacobian, rows stay zero
    obs2 = _observer(c, 2, "unit", rng, bkg_T=None)
    got, ref = _gpu(wsm, c, TARGETS, obs2), _oracle(orc, c, TARGETS, obs2)
    _compare(got, ref)
    assert not got[1][-1].any() and not ref[1][-1].any()
    # forward only: no state vector at all
    obs3 = _observer(c, 0, "PlanckBT", rng, surface_rows=False)
    obs3.nx = 0
    got, ref = _gpu(wsm, c, (), obs3), _oracle(orc, c, (), obs3)
    np.testing.assert_allclose(got[0], ref[0], rtol=1e-9)
    np.testing.assert_allclose(got[2], ref[2], rtol=1e-9)


def test_observer_matches_separate_path_jacobian(wsm):
    """The x-space accumulation inside the pass equals W^T dI formed from the downloaded per-level Jacobian."""
    c = synth.tiny_case(nl=64, nf=300, np_=7, targets=TARGETS)
    obs = _observer(c, 2, "unit", np.random.default_rng(5), bkg_T=None)
    obs.n_real = 1.0
    cat = wsm.Catalog(c.cat)
    path = wsm.Path(cat, c.nf, c.np_, 2)
    path.upload(c.f, c.atm, c.r, c.I_bkg, targets=TARGETS)
    path.run_propmat()
    path.run_stokes()
    I = np.empty((c.nf, 4)); dI = np.empty((c.nf, c.np_, 2, 4))
    path.download(I=I, dI=dI)
    path.run_observer(obs)
    Io, Jx, y, Jy = path.download_observer()
    path.close(); cat.close()
    assert np.array_equal(Io, I)
    ref = np.zeros_like(Jx)
    for ip in range(c.np_):
        for t in range(2):
            for (x, w) in obs.path_map[ip][t]:
                ref[x] += w * dI[:, ip, t]
    np.testing.assert_allclose(Jx, ref, rtol=1e-12, atol=1e-15 * np.abs(ref).max())


def test_observer_rejects_bad_input(wsm):
    c = synth.tiny_case(nl=20, nf=64, np_=3, targets=TARGETS)
    cat = wsm.Catalog(c.cat)
    path = wsm.Path(cat, c.nf, c.np_, 2)
    obs = _observer(c, 2, "unit", np.random.default_rng(6))
    with pytest.raises(wsm.Ab200Error):  # not uploaded
        path.run_observer(obs)
    path.upload(np.tile(c.f, (c.np_, 1)), c.atm, c.r, c.I_bkg, targets=TARGETS)
    path.run_propmat()
    with pytest.raises(wsm.Ab200Error, match="sensor's frequency grid"):
        path.run_observer(obs)
    path.upload(c.f, c.atm, c.r, c.I_bkg, targets=TARGETS)
    path.run_propmat()
    bad = _observer(c, 2, "unit", np.random.default_rng(6))
    bad.path_map[1][0][0] = (bad.nx, 0.5)
    with pytest.raises(wsm.Ab200Error, match="outside the state vector"):
        path.run_observer(bad)
    bad2 = _observer(c, 2, "unit", np.random.default_rng(6))
    bad2.channels[0][0] = (c.nf, (1.0, 0, 0, 0))
    with pytest.raises(wsm.Ab200Error, match="outside the frequency grid"):
        path.run_observer(bad2)
    path.run_observer(obs)  # still usable afterwards
    path.download_observer()
    path.close(); cat.close()


def test_measurement_vec_from_sensor_ragged_batch(wsm, orc):
    """measurement_vecFromSensor (src/m_rad.cc:301-362): a batch of paths of different lengths, one workspace pair,
    only the channel contributions leave the device; against the sum of the oracle's per-path results."""
    rng = np.random.default_rng(8)
    base = synth.tiny_case(nl=64, nf=200, np_=7, targets=TARGETS)
    sims, yref, Jref = [], 0.0, 0.0
    for k, np_ in enumerate((7, 4, 6, 5, 7)):
        c = synth.tiny_case(nl=64, nf=200, np_=np_, targets=TARGETS)
        c.atm.T = c.atm.T + 0.7 * k  # different paths see different atmospheres
        obs = _observer(c, 2, "PlanckBT", rng, bkg_T=280.0 + k)
        sims.append((c.atm, c.r, obs))
        _, _, y, Jy = _oracle(orc, c, TARGETS, obs)
        yref, Jref = yref + y, Jref + Jy
    y, J = wsm.measurement_vecFromSensor(base.cat, base.f, sims, jac_targets=TARGETS)
    np.testing.assert_allclose(y, yref, rtol=1e-9)
    assert np.abs(J - Jref).max() <= 2e-7 * np.abs(Jref).max()
    # one workspace, same numbers bit for bit (the order of accumulation is the order of the simulations)
    y1, J1 = wsm.measurement_vecFromSensor(base.cat, base.f, sims, jac_targets=TARGETS, n_workspaces=1)
    assert np.array_equal(y, y1) and np.array_equal(J, J1)


def test_observer_with_wind_magnetic_and_line_targets(wsm, orc):
    """The state-space accumulation and the sensor sum-up are target-agnostic: wind, magnetic-field and line-parameter rows
    (scalar dK rows, polarised dK rows, rows of one line) go through the same epilogue as T / VMR."""
    c = synth.tiny_case(nf=38 * 6, np_=5, zeeman=True, targets=TARGETS)
    rng = np.random.default_rng(7)  # a calm path: its wind rows are not zero
    line = int(np.flatnonzero(c.cat.z_on)[0])
    tg = (("T",), ("wind_w",), ("mag_v",), ("line_a", line), ("VMR", 0), ("line_ls", line, abi.VAR_G0, abi.SPECIES_BATH, 0))
    # RJBT: the Planck brightness-temperature transform scales a Q / U / V row by dinvplanckdI((I + V) / 2) -
    # dinvplanckdI((I - V) / 2) (spectral_radiance_transform_operator.cc:46-87), a difference of nearly equal numbers
    # where |V| / I ~ 1e-8 as around this fixture's 118 GHz line: there a 1e-12 difference in I shows up at 1e-5 in the row
    obs = _observer(c, len(tg), "RJBT", rng, bkg_T=260.0)
    ref = _oracle(orc, c, tg, obs)
    got = _gpu(wsm, c, tg, obs)
    _compare(got, ref, jac_rtol=5e-7)
    n_grid = 4
    for t in range(len(tg)):
        assert np.abs(ref[1][t * n_grid:(t + 1) * n_grid]).max() > 0, tg[t]
