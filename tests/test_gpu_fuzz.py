"""Randomised parity sweep of the whole path against the oracle: random catalogs (every temperature model, with and
without Bath broadener, line mixing, Zeeman lines, ByLine cutoffs, several bands per species), random atmospheres
from the surface to the mesosphere, frequency grids that resolve line cores and reach far wings, random targets."""
import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import synth
from arts_b200._abi import AtmPath, HostCatalog
from tests.conftest import assert_propmat_close

pytestmark = pytest.mark.gpu


def random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    n_species = int(rng.integers(1, 4))
    n_bands = int(rng.integers(1, 6))
    band_isot = rng.integers(0, n_species, n_bands).astype(np.int32)
    sizes = rng.integers(1, 140, n_bands)
    band_offset = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    nl = int(band_offset[-1])
    f_lo, f_hi = (50e9, 70e9) if seed % 3 else (300e9, 2e12)
    f0 = np.concatenate([np.sort(rng.uniform(f_lo, f_hi, s)) for s in sizes])
    zeeman_band = rng.random(n_bands) < 0.3
    cutoff_band = rng.random(n_bands) < 0.3
    mixing_band = rng.random(n_bands) < 0.3
    band_of = np.repeat(np.arange(n_bands), sizes)
    a = 4.479e-9 * (f0 / 118.75e9) ** 3 * 10 ** rng.uniform(-2, 1, nl)
    e0 = 1e-23 * rng.uniform(0.5, 60, nl)
    Jl = rng.integers(0, 6, nl)
    Ju = np.maximum(Jl + rng.integers(-1, 2, nl), 0)
    nb = int(rng.integers(1, 3))  # broadeners per line
    with_bath = bool(rng.integers(0, 2)) or nb == 1
    n_ls = nl * nb
    ls_offset = np.arange(nl + 1, dtype=np.int64) * nb
    ls_species = np.empty(n_ls, np.int32)
    for k in range(nb):
        ls_species[k::nb] = abi.SPECIES_BATH if (with_bath and k == nb - 1) else rng.integers(0, n_species)
    if not with_bath and nb == 2:  # two distinct non-bath broadeners
        ls_species[0::2] = 0
        ls_species[1::2] = (n_species - 1) if n_species > 1 else 0
    ls_type = np.full((n_ls, abi.NVAR), abi.TM_ABSENT, np.int32)
    ls_X = np.zeros((n_ls, abi.NVAR, 4))
    models = [abi.TM_T0, abi.TM_T1, abi.TM_T2, abi.TM_T3, abi.TM_T4, abi.TM_T5, abi.TM_AER, abi.TM_DPL, abi.TM_POLY]
    g_type = rng.choice(models, n_ls)
    ls_type[:, abi.VAR_G0] = g_type
    g = rng.uniform(0.8e4, 3e4, n_ls)
    ex = rng.uniform(0.5, 0.9, n_ls)
    X = ls_X[:, abi.VAR_G0]
    X[:, 0] = g
    X[:, 1] = ex
    for i, t in enumerate(g_type):  # keep every model's value ~g, positive, over 180-300 K
        if t == abi.TM_T2: X[i, 2] = rng.uniform(-0.1, 0.1)
        elif t == abi.TM_T3: X[i, 1] = -g[i] * 1e-3
        elif t == abi.TM_T4: X[i, 1], X[i, 2] = g[i] * 0.1, ex[i]
        elif t == abi.TM_T5: X[i, 1] = rng.uniform(0.2, 0.4)
        elif t == abi.TM_AER: X[i, :] = g[i] * np.array([1.3, 1.15, 1.0, 0.9])
        elif t == abi.TM_DPL: X[i, 2], X[i, 3] = g[i] * 0.1, 0.3
        elif t == abi.TM_POLY: X[i, :] = [g[i] * 1.5, -g[i] * 2e-3, 0.0, 0.0]
    ls_type[:, abi.VAR_D0] = abi.TM_T1
    ls_X[:, abi.VAR_D0, 0] = rng.uniform(-800, 800, n_ls)
    ls_X[:, abi.VAR_D0, 1] = rng.uniform(0.5, 1.0, n_ls)
    mix_ls = np.repeat(mixing_band[band_of], nb)
    ls_type[mix_ls, abi.VAR_Y] = abi.TM_T1
    ls_X[mix_ls, abi.VAR_Y, 0] = rng.uniform(-3e-6, 3e-6, mix_ls.sum())
    ls_X[mix_ls, abi.VAR_Y, 1] = 0.8
    ls_type[mix_ls, abi.VAR_G] = abi.TM_T1
    ls_X[mix_ls, abi.VAR_G, 0] = rng.uniform(-1e-11, 1e-11, mix_ls.sum())
    ls_X[mix_ls, abi.VAR_G, 1] = 0.5
    cat = HostCatalog(
        n_species=n_species, isot_species=np.arange(n_species), isot_mass=rng.uniform(18, 48, n_species), band_isot=band_isot,
        band_offset=band_offset, f0=f0, a=a, e0=e0, gu=2.0 * Ju + 1, gl=2.0 * Jl + 1, T0=np.full(nl, 296.0),
        ls_offset=ls_offset, ls_species=ls_species, ls_type=ls_type, ls_X=ls_X,
        band_cutoff_type=np.where(cutoff_band, abi.CUTOFF_BYLINE, abi.CUTOFF_NONE),
        band_cutoff_value=np.where(cutoff_band, rng.uniform(0.05, 3.0, n_bands) * (f_hi - f_lo) / 10, np.inf),
        z_on=zeeman_band[band_of].astype(np.uint8), z_gu=rng.uniform(0.5, 2.0, nl), z_gl=rng.uniform(0.5, 2.0, nl),
        two_Ju=2 * Ju, two_Jl=2 * Jl)
    np_ = int(rng.integers(1, 6))
    z = np.sort(rng.uniform(0, 90, np_))[::-1]
    T, P = synth.standard_profile(z)
    vmr = rng.uniform(1e-4, 0.3, (np_, n_species))
    Q = (rng.uniform(100, 400, n_species)[None, :] * (T / 296.0)[:, None])
    atm = AtmPath(T=T, P=P, vmr=vmr, isorat=rng.uniform(0.9, 1.0, (np_, n_species)), Q=Q, dQdT=Q / T[:, None],
                  mag=rng.normal(0, 30e-6, (np_, 3)), los=np.tile([rng.uniform(95, 180), rng.uniform(-180, 180)], (np_, 1)),
                  wind=rng.normal(0, 40, (np_, 3)) if seed % 4 == 0 else None)
    # grid: broadband + dense clusters around a few line centres (cores, CF / series regions)
    centres = rng.choice(f0, min(4, nl), replace=False)
    f = np.unique(np.concatenate([np.linspace(f_lo, f_hi, int(rng.integers(40, 500)))] +
                                 [c + rng.uniform(-1, 1, 60) * 10 ** rng.uniform(3, 8) for c in centres]))
    r = np.abs(np.diff(z)) * 1e3 / abs(np.cos(np.deg2rad(atm.los[0, 0])))
    bkg = np.zeros((len(f), 4))
    bkg[:, 0] = synth.planck(f, 2.735)
    return cat, f, atm, r, bkg


@pytest.mark.parametrize("seed", range(24))
def test_random_case(wsm, orc, seed):
    cat, f, atm, r, bkg = random_case(seed)
    clamp = seed % 2
    Kr, _ = orc.propmat_levels(cat, f, atm, no_negative_absorption=clamp)
    K, _ = wsm.spectral_propmat_pathFromPath(cat, f, atm, no_negative_absorption=clamp)
    assert_propmat_close(K, Kr, atol_scale=1e-10, what=f"seed {seed} K")
    if atm.np_ < 2:
        return
    option = ("linsrc", "constant")[seed % 2]
    tg = ((("T",), ("VMR", 0)) if seed % 3 == 0 else ())
    Ir, dIr = orc.clearsky_emission(cat, f, atm, r, bkg, rte_option=option, no_negative_absorption=clamp, targets=tg,
                                    hse_derivative=seed % 2)
    I, dI = wsm.spectral_radClearskyEmission(cat, f, atm, r, bkg, rte_option=option, no_negative_absorption=clamp,
                                             jac_targets=tg, hse_derivative=seed % 2)
    tb, tbr = wsm.spectral_radApplyPlanckTb(I, f), orc.planck_tb(f, Ir)
    assert np.abs(tb - tbr).max() <= 1e-6, f"seed {seed} Tb"
    for q in range(len(tg)):
        sc = np.abs(dIr[:, :, q]).reshape(-1, 4).max(axis=0)
        err = np.abs(dI[:, :, q] - dIr[:, :, q]).reshape(-1, 4).max(axis=0)
        assert (err <= 2e-6 * sc + 1e-300).all(), f"seed {seed} dI target {q}: {err / np.maximum(sc, 1e-300)}"
