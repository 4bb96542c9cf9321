"""GPU parity of the Jacobian rows of the path (a10, a15, a16, a19) against the CPU oracle:
d(propmat)/d(T | VMR) with the reference's forward-finite-difference dF (lbl_lineshape_voigt_lte.cpp:250-268),
tran::deriv / linsrc_deriv (rtepack_transmission.cc:277-447,558-674) and the dI accumulation of rte_emission
(rtepack_rtestep.cc:293-306,348-367), un-fused and fused.

Tolerance: the reference's dF divides a difference of two w(z) values by 1e-4 |z|, which amplifies the
~1e-13 accuracy of ANY w(z) implementation (the reference's own included) by 1e4; Jacobians are therefore
compared at 2e-7 of the largest element of each component instead of the 1e-9 of the forward model."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from tests.conftest import assert_propmat_close

pytestmark = pytest.mark.gpu


def assert_jac_close(a, b, rtol=2e-7, what="jacobian"):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.isfinite(a).all(), f"{what}: non-finite"
    scale = np.abs(b).reshape(-1, b.shape[-1]).max(axis=0)
    err = np.abs(a - b) / np.maximum(scale, 1e-300)
    bad = (err > rtol) & (np.abs(a - b) > 0)
    assert not bad.any(), f"{what}: worst {err.max():.3e} at {np.unravel_index(np.argmax(err), err.shape)}"


TARGETS = (("T",), ("VMR", 0), ("VMR", 1))


def test_propmat_jacobian_scalar(wsm, orc):
    c = synth.tiny_case(nl=64, nf=300, np_=4)
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=TARGETS)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=TARGETS)
    assert dK.shape == (4, 3, 300, 7)
    for q in range(3):
        assert_jac_close(dK[:, q], dKr[:, q], what=f"dK target {q}")
    assert np.abs(dKr).max() > 0
    # the forward part is unchanged by the presence of targets (bitwise)
    K0, _ = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm)
    assert np.array_equal(K, K0)


def test_propmat_jacobian_more_than_one_pass(wsm, orc):
    """6 targets = two passes of 4 in the kernel."""
    c = synth.case_c2(lines_per_species=40, nf=200, np_=3, bands_per_species=2)
    tg = (("VMR", 0), ("VMR", 1), ("T",), ("VMR", 2), ("VMR", 3), ("VMR", 4))
    _, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    _, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    for q in range(6):
        assert_jac_close(dK[:, q], dKr[:, q], what=f"dK target {q}")


@pytest.mark.parametrize("cutoff", [None, 2e9])
def test_propmat_jacobian_cutoff_and_mixing(wsm, orc, cutoff):
    c = synth.case_c1(nl=60, nf=500, cutoff=cutoff)
    c.cat.ls_type[:, abi.VAR_Y] = abi.TM_T1
    rng = np.random.default_rng(5)
    n = len(c.cat.ls_species)
    c.cat.ls_X[:, abi.VAR_Y, 0] = rng.uniform(-3e-6, 3e-6, n)
    c.cat.ls_X[:, abi.VAR_Y, 1] = 0.8
    c.cat.ls_type[:, abi.VAR_DV] = abi.TM_T1
    c.cat.ls_X[:, abi.VAR_DV, 0] = rng.uniform(-1e-3, 1e-3, n)
    c.cat.ls_X[:, abi.VAR_DV, 1] = 0.6
    tg = (("T",), ("VMR", 0))
    _, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg, no_negative_absorption=0)
    _, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg, no_negative_absorption=0)
    for q in range(2):
        assert_jac_close(dK[:, q], dKr[:, q], what=f"dK target {q} cutoff {cutoff}")


def test_propmat_jacobian_zeeman(wsm, orc):
    c = synth.case_c3(nf=38 * 12, np_=3, los=(120.0, 30.0))
    tg = (("T",), ("VMR", 0))
    _, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    _, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    for q in range(2):
        assert_jac_close(dK[:, q], dKr[:, q], what=f"zeeman dK target {q}")
    assert np.abs(dK[..., 4:]).max() > 0


def test_addlines_accumulates_jacobian(wsm, orc):
    c = synth.case_c1(nl=30, nf=111)
    tg = (("T",),)
    _, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K = np.zeros((c.nf, 7))
    dK = np.full((1, c.nf, 7), 0.5)
    wsm.spectral_propmatAddLines(K, dK, c.f, tg, abi.SPECIES_BATH, c.cat, c.atm)
    scale = np.abs(dKr).max()
    assert np.abs((dK - 0.5) - dKr[0]).max() <= 1e-6 * scale + 1e-15


def _random_K(rng, np_, nf, polarised, scale=1e-4):
    K = np.zeros((np_, nf, 7))
    K[..., 0] = rng.uniform(0.2, 2.0, (np_, nf)) * scale
    if polarised:
        K[..., 1:] = rng.uniform(-0.3, 0.3, (np_, nf, 6)) * scale
        K[:, ::5, 1:] = 0.0
    return K


@pytest.mark.parametrize("polarised", [False, True])
@pytest.mark.parametrize("option", ["constant", "linsrc"])
def test_unfused_jacobian_chain(wsm, orc, polarised, option):
    rng = np.random.default_rng(4)
    np_, nf, nq = 7, 150, 2
    K = _random_K(rng, np_, nf, polarised)
    dK = np.zeros((np_, nq, nf, 7))
    dK[..., 0] = rng.uniform(-1, 1, (np_, nq, nf)) * 1e-6
    if polarised:
        dK[..., 1:] = rng.uniform(-1, 1, (np_, nq, nf, 6)) * 1e-7
    r = rng.uniform(200.0, 3000.0, np_ - 1)
    Tlev = np.linspace(210.0, 290.0, np_)
    f = np.linspace(50e9, 70e9, nf)
    bkg = np.zeros((nf, 4))
    bkg[:, 0] = synth.planck(f, 2.735)
    dr = np.zeros((2, np_ - 1, nq))
    dr[0, :, 0] = r / (2 * Tlev[:-1])
    dr[1, :, 0] = r / (2 * Tlev[1:])

    Tr, Lr, Pr, dTr, dLr = orc.tramat(K, dK, r, dr, option)
    tm = wsm.spectral_tramat_pathFromPath(K, dK, r, Tlev, option, hse_derivative=1, it=0)
    np.testing.assert_allclose(tm.T.reshape(Tr.shape), Tr, rtol=1e-11, atol=1e-13)
    sc = np.abs(dTr).max()
    np.testing.assert_allclose(tm.dT.reshape(dTr.shape), dTr, rtol=1e-9, atol=1e-11 * sc, err_msg="dT")
    if option == "linsrc":
        sc = np.abs(dLr).max()
        np.testing.assert_allclose(tm.dL.reshape(dLr.shape), dLr, rtol=1e-8, atol=1e-10 * sc, err_msg="dL")
    Jr, dJr = orc.srcvec(K, f, Tlev, it=0, nq=nq)
    J, dJ = wsm.spectral_rad_srcvec_pathFromPropmat(K, f, Tlev, it=0, nq=nq)
    np.testing.assert_allclose(dJ, dJr, rtol=1e-13)
    Ir, dIr = orc.rte_emission(option, Tr, Lr, Pr, dTr, dLr, Jr, dJr, bkg)
    I, dI = wsm.spectral_radStepByStepEmission(tm, J, dJ, bkg)
    np.testing.assert_allclose(I, Ir, rtol=1e-11, atol=1e-13 * np.abs(Ir).max())
    np.testing.assert_allclose(dI, dIr, rtol=1e-8, atol=1e-10 * np.abs(dIr).max(), err_msg="dI")


@pytest.mark.parametrize("option", ["constant", "linsrc"])
@pytest.mark.parametrize("hse", [0, 1])
def test_fused_clearsky_jacobian_scalar(wsm, orc, option, hse):
    """Reduced BASELINE config 5 (one path): T and VMR Jacobians of the radiance."""
    c = synth.case_c5_single(n_lines=500, nf=600, np_=15)
    tg = (("T",), ("VMR", 0), ("VMR", 3))
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option, targets=tg, hse_derivative=hse)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option, jac_targets=tg,
                                             hse_derivative=hse)
    np.testing.assert_allclose(I, Ir, rtol=1e-9)
    assert dI.shape == (600, 15, 3, 4)
    for q in range(3):
        assert_jac_close(dI[:, :, q, :1], dIr[:, :, q, :1], rtol=5e-7, what=f"dI target {q}")
    assert np.abs(dIr[..., 0]).max() > 0
    # forward result identical with and without targets
    I0, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option)
    assert np.array_equal(I, I0)


@pytest.mark.parametrize("option", ["constant", "linsrc"])
def test_fused_clearsky_jacobian_polarised(wsm, orc, option):
    c = synth.case_c3(nf=38 * 6, np_=8, los=(120.0, 30.0), rte_option=option)
    tg = (("T",), ("VMR", 0))
    Ir, dIr = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option, targets=tg, hse_derivative=1)
    I, dI = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=option, jac_targets=tg,
                                             hse_derivative=1)
    np.testing.assert_allclose(I, Ir, rtol=1e-9, atol=1e-12 * np.abs(Ir).max())
    for q in range(2):
        assert_jac_close(dI[:, :, q], dIr[:, :, q], rtol=5e-7, what=f"polarised dI target {q}")


def test_jacobian_errors_are_loud(wsm):
    from arts_b200._lib import Ab200Error

    c = synth.tiny_case(nl=16, nf=32, np_=3)
    with pytest.raises(Ab200Error) as e:
        wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=(("VMR", 7),))
    assert e.value.code == abi.ERR_INVALID
    with pytest.raises(Ab200Error) as e:
        wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=(("T",),), flags=abi.FLAG_TRAN_EXACT)
    assert e.value.code == abi.ERR_UNSUPPORTED
    with pytest.raises(Ab200Error) as e:
        wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=((99, 0),))
    assert e.value.code == abi.ERR_UNSUPPORTED


def test_reentrant_from_host_threads(wsm, orc):
    """The shims are called concurrently from OpenMP worker threads (src/m_rad.cc:321-343): every entry point
    must be re-entrant with one device workspace per host thread and a shared immutable catalog."""
    import threading

    cases = [synth.case_c5_single(n_lines=200, nf=300 + 37 * i, np_=10 + i) for i in range(4)]
    cat = wsm.Catalog(cases[0].cat)  # same catalog for all paths
    tg = (("T",), ("VMR", 0))
    out, err = [None] * len(cases), []

    def work(i):
        try:
            for _ in range(3):
                c = cases[i]
                out[i] = wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg)
            wsm.lib().ab200_release_thread_cache()
        except Exception as e:  # noqa: BLE001
            err.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    for i, c in enumerate(cases):
        I, dI = wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg)
        assert np.array_equal(out[i][0], I) and np.array_equal(out[i][1], dI)
    cat.close()


def test_propmat_jacobian_far_tiles_without_bath_broadener(wsm, orc):
    """Without a bath broadener the reference's d/dVMR of an ABSENT Y or G model is the sum of the broadeners' VMRs
    (lbl_lineshape_model.cpp:112), so the strength derivative of a plain Voigt line has an imaginary part and
    Re(ds F) picks up -Im ds Im F, which dominates in the far wings.  Many lines, so that whole tiles are far from
    whole frequency blocks and the closed-form tile path is the one that runs."""
    c = synth.case_c2(lines_per_species=1200, nf=1500, np_=3, bands_per_species=3)
    bath = c.cat.ls_species == abi.SPECIES_BATH
    c.cat.ls_species[bath] = 4  # every line: self + species 4, no bath
    tg = (("VMR", 0), ("VMR", 4), ("T",))
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    for q in range(3):
        # per level: the far wings of the upper levels must not hide behind the line cores of the lowest one
        for lev in range(c.np_):
            assert_jac_close(dK[lev, q], dKr[lev, q], what=f"dK target {q} level {lev}")
    assert np.abs(dKr[:, 0]).max() > 0


@pytest.mark.parametrize("cutoff", [None, 2e9])
def test_isotopologue_ratio_rows(wsm, orc, cutoff):
    """compute_derivative(SpeciesIsotope), lbl_lineshape_voigt_lte.cpp:1526-1544: as a derivative record (ds = s / ratio)
    inside species-merged segments, with cutoffs, next to other targets, and for Zeeman bands."""
    c = synth.tiny_case(nl=300, nf=700, np_=4, cutoff=cutoff)
    ni = len(c.cat.isot_species)
    tg = [("isorat", 0), ("T",), ("isorat", ni - 1), ("VMR", 0), ("isorat", 1 % ni)]
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    assert_propmat_close(K, Kr, atol_scale=1e-11 if cutoff else 1e-12)
    for q in range(len(tg)):
        assert np.abs(dKr[:, q]).max() > 0
        assert_jac_close(dK[:, q], dKr[:, q], what=f"dK target {tg[q]}")
    z = synth.tiny_case(nf=38 * 8, np_=3, zeeman=True)
    Kr, dKr = orc.propmat_levels(z.cat, z.f, z.atm, targets=[("isorat", 0)])
    K, dK = wsm.spectral_propmat_pathFromPath(z.cat, z.f, z.atm, jac_targets=[("isorat", 0)])
    assert_jac_close(dK[:, 0], dKr[:, 0], rtol=1e-9, what="Zeeman isotopologue-ratio row")
    np.testing.assert_allclose(dK[:, 0] * z.atm.isorat[:, 0][:, None, None], K, rtol=1e-9, atol=1e-12 * np.abs(K).max())
    z.atm.isorat[1, 0] = 0.0
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.spectral_propmat_pathFromPath(z.cat, z.f, z.atm, jac_targets=[("isorat", 0)])
    assert e.value.code == abi.ERR_INVALID and "isotopologue ratios" in str(e.value)


def _many_lines_case(np_=3, nf=2600):
    """Enough lines and grid that every tile class occurs for a 512-frequency block: wholly very far (|x| > 1.2e4 for every
    pair), mixed (the pair test decides, both signs of f - f0') and near."""
    c = synth.case_c2(lines_per_species=900, nf=nf, np_=np_, bands_per_species=3)
    return c


def test_very_far_pairs_closed_form(wsm, orc):
    """Pairs with |x| > 1.2e4 of real lines take ONE rational function shared by all targets (lbl_sum_jac_vfar_kernel):
    F to 1e-12, the forward difference dF to 1 / |z|^2 <= 6.9e-9 per pair.  Against the oracle's literal forward difference, per level (the far
    wings of the upper levels must not hide behind the line cores of the lowest one), T + own + foreign VMR + a wind row."""
    c = _many_lines_case()
    tg = (("T",), ("VMR", 0), ("VMR", 3), ("wind_u",))
    c.atm.wind = np.tile(np.array([12.0, -7.0, 0.4]), (c.np_, 1))
    c.atm.los = np.tile(np.array([137.0, 21.0]), (c.np_, 1))
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    assert_propmat_close(K, Kr)
    for q in range(len(tg)):
        for lev in range(c.np_):
            assert np.abs(dKr[lev, q]).max() > 0
            assert_jac_close(dK[lev, q], dKr[lev, q], rtol=2e-8, what=f"dK target {tg[q]} level {lev}")


def test_jacobian_rows_do_not_depend_on_the_frequency_partition(wsm):
    """Which closed form a (line, frequency) pair takes is a property of the pair alone, so a Jacobian row at a frequency
    has the same bits on the whole grid and on any sub-grid (shard), whatever the block boundaries."""
    c = _many_lines_case(np_=2, nf=3000)
    tg = (("T",), ("VMR", 1))
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    for lo, hi in ((0, 700), (333, 1900), (1777, 3000)):
        Ks, dKs = wsm.spectral_propmat_pathFromPath(c.cat, c.f[lo:hi], c.atm, jac_targets=tg)
        assert np.array_equal(Ks, K[:, lo:hi]), (lo, hi)
        assert np.array_equal(dKs, dK[:, :, lo:hi]), (lo, hi)


def test_very_far_pairs_complex_lines(wsm, orc):
    """Zeeman components and line mixing (complex strengths): the very far pairs of the configs[2] band — every other line of
    the band is 0.5 GHz or more away in units of a 70 kHz Doppler width — take the complex one-denominator form
    (lbl_sum_jac_vfar_cplx_kernel).  T, VMR, the three magnetic-field components and a wind row, per level."""
    c = synth.case_c3(nf=38 * 24, np_=4, los=(120.0, 30.0))
    tg = (("T",), ("VMR", 0), ("mag_u",), ("mag_v",), ("mag_w",), ("wind_v",))
    c.atm.wind = np.tile(np.array([5.0, -3.0, 0.2]), (c.np_, 1))
    Kr, dKr = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
    assert_propmat_close(K, Kr)
    for q in range(len(tg)):
        for lev in range(c.np_):
            assert np.abs(dKr[lev, q]).max() > 0
            assert_jac_close(dK[lev, q], dKr[lev, q], rtol=5e-8, what=f"dK target {tg[q]} level {lev}")
    # sub-grid: the pair rule is the same; the other (near) classes of complex lines are block-dependent in the last bits
    lo, hi = 100, 700
    Ks, dKs = wsm.spectral_propmat_pathFromPath(c.cat, c.f[lo:hi], c.atm, jac_targets=tg)
    for q in range(len(tg)):
        assert_jac_close(dKs[:, q], dK[:, q, lo:hi], rtol=1e-12, what=f"sub-grid row {tg[q]}")
