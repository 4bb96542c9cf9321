"""Synthetic catalogs / atmospheres for the BASELINE.json configs.

All generators are seeded (numpy ``default_rng``) and generalise the reference's own
synthetic-line recipe (src/core/lbl/test/test_lbl_perf.cpp:29-53: O2-66-like lines,
``a = 4.479e-9 (1+U)``, ``e0 = 1e-23 (1+U)``, ``gu = 3``, T1 broadening, T0 = 296 K).
Partition functions are synthetic (``Q(T) = Q0 T/296``) because the reference generates
its real ones from external data at build time (src/partfun/CMakeLists.txt:8-24).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi as abi
from ._abi import AtmPath, HostCatalog

H = 6.62607015e-34
KB = 1.380649e-23
C0 = 299792458.0
T_CMB = 2.735  # arts_constants.h:283


def planck(f, T):
    """physics_funcs.cc:192-197 (numpy, for backgrounds only)."""
    a = 2 * H / C0**2
    b = H / KB
    return a * f**3 / np.expm1(b * f / T)


@dataclass
class Case:
    """One synthetic clear-sky case: everything the WSM chain needs."""

    name: str
    cat: HostCatalog
    f: np.ndarray  # [nf] ascending frequency grid (shared by all levels)
    atm: AtmPath
    r: np.ndarray  # [np-1] layer lengths [m]
    I_bkg: np.ndarray  # [nf,4]
    rte_option: str = "linsrc"
    select_species: int = abi.SPECIES_BATH
    no_negative_absorption: int = 1
    targets: tuple = ()
    note: str = ""

    @property
    def nf(self):
        return len(self.f)

    @property
    def np_(self):
        return self.atm.np_

    @property
    def n_lines(self):
        return self.cat.n_lines


def _ls_tables(nl, rng, species_self, g0=(1e4, 3e4), g0_x1=(0.5, 1.0), d0=None, y=None, with_bath=True):
    """Two broadeners per line (self + Bath, i.e. 'AIR'-style) with T1 models."""
    nb = 2 if with_bath else 1
    n_ls = nl * nb
    ls_offset = np.arange(nl + 1, dtype=np.int64) * nb
    ls_species = np.empty(n_ls, np.int32)
    ls_species[0::nb] = species_self
    if with_bath:
        ls_species[1::nb] = abi.SPECIES_BATH
    ls_type = np.full((n_ls, abi.NVAR), abi.TM_ABSENT, np.int32)
    ls_X = np.zeros((n_ls, abi.NVAR, 4))
    ls_type[:, abi.VAR_G0] = abi.TM_T1
    ls_X[:, abi.VAR_G0, 0] = rng.uniform(*g0, n_ls)
    ls_X[:, abi.VAR_G0, 1] = rng.uniform(*g0_x1, n_ls)
    if d0 is not None:
        ls_type[:, abi.VAR_D0] = abi.TM_T1
        ls_X[:, abi.VAR_D0, 0] = rng.uniform(*d0, n_ls)
        ls_X[:, abi.VAR_D0, 1] = rng.uniform(0.5, 1.0, n_ls)
    if y is not None:
        ls_type[:, abi.VAR_Y] = abi.TM_T1
        ls_X[:, abi.VAR_Y, 0] = rng.uniform(*y, n_ls)
        ls_X[:, abi.VAR_Y, 1] = 0.8
    return ls_offset, ls_species, ls_type, ls_X


def standard_profile(z_km):
    """T(z) piecewise linear 288 -> 217 -> 217 -> 271 -> 271 -> 190 K, P = 1013.25 hPa exp(-z/7 km)."""
    zk = np.array([0.0, 11.0, 20.0, 47.0, 51.0, 85.0, 120.0])
    tk = np.array([288.0, 217.0, 217.0, 271.0, 271.0, 190.0, 190.0])
    T = np.interp(z_km, zk, tk)
    P = 101325.0 * np.exp(-z_km / 7.0)
    return T, P


# ---------------------------------------------------------------------------
def case_c1(nl=1000, nf=10_000, seed=1, cutoff=None) -> Case:
    """C1: one species, 1k Voigt lines, 1e4 frequencies, one level, propmat only."""
    rng = np.random.default_rng(seed)
    f0 = np.sort(rng.uniform(100e9, 130e9, nl))
    a = 4.479289583303983e-09 * (1 + rng.uniform(0, 1, nl))
    e0 = 1e-23 * (1 + rng.uniform(0, 1, nl))
    ls_offset, ls_species, ls_type, ls_X = _ls_tables(nl, rng, 0, d0=(-500.0, 500.0))
    cat = HostCatalog(
        n_species=1, isot_species=[0], isot_mass=[31.9898], band_isot=[0], band_offset=[0, nl],
        f0=f0, a=a, e0=e0, gu=np.full(nl, 3.0), gl=np.full(nl, 1.0), T0=np.full(nl, 296.0),
        ls_offset=ls_offset, ls_species=ls_species, ls_type=ls_type, ls_X=ls_X,
        band_cutoff_type=None if cutoff is None else [abi.CUTOFF_BYLINE],
        band_cutoff_value=None if cutoff is None else [cutoff],
    )
    T = np.array([250.0])
    atm = AtmPath(T=T, P=[1e4], vmr=[[0.21]], isorat=[[0.995]], Q=[[215.0 * 250.0 / 296.0]], dQdT=[[215.0 / 296.0]])
    f = np.linspace(100e9, 130e9, nf)
    return Case("C1", cat, f, atm, r=np.zeros(0), I_bkg=np.zeros((nf, 4)), note="propmat only")


# ---------------------------------------------------------------------------
_C2_MASS = [18.01, 43.99, 47.98, 31.99, 28.01]
_C2_VMR0 = [1e-2, 4e-4, 1e-6, 0.21, 1e-7]
# strength scales chosen once (by running the oracle on a coarse grid) so that the nadir opacity of the
# synthetic atmosphere spans ~1e-2..1e2 over the 1-1000 GHz grid
_C2_STRENGTH = [6e-3, 0.2, 60.0, 4e-4, 600.0]


def _multi_species_catalog(rng, n_species, lines_per_species, bands_per_species, f_lo, f_hi, masses, strength,
                           decades=3.0, cutoff=None):
    nl = n_species * lines_per_species
    per_band = lines_per_species // bands_per_species
    assert per_band * bands_per_species == lines_per_species
    nb = n_species * bands_per_species
    f0 = np.empty(nl)
    band_isot = np.repeat(np.arange(n_species, dtype=np.int32), bands_per_species)
    band_offset = np.arange(nb + 1, dtype=np.int64) * per_band
    for b in range(nb):
        f0[b * per_band:(b + 1) * per_band] = np.sort(rng.uniform(f_lo, f_hi, per_band))
    # a ~ f0^3 keeps the LTE strength a*gu/f0^3 (lbl_data.h:66-68) frequency independent
    sp_of_line = np.repeat(np.arange(n_species), lines_per_species)
    a = (4.479289583303983e-09 * (f0 / 118.750348e9) ** 3 * 10.0 ** (-decades * rng.uniform(0, 1, nl))
         * np.asarray(strength)[sp_of_line])
    e0 = 1e-23 * (1 + 40 * rng.uniform(0, 1, nl))
    ls_offset = np.arange(nl + 1, dtype=np.int64) * 2
    ls_species = np.empty(2 * nl, np.int32)
    ls_species[0::2] = sp_of_line
    ls_species[1::2] = abi.SPECIES_BATH
    ls_type = np.full((2 * nl, abi.NVAR), abi.TM_ABSENT, np.int32)
    ls_X = np.zeros((2 * nl, abi.NVAR, 4))
    ls_type[:, abi.VAR_G0] = abi.TM_T1
    ls_X[:, abi.VAR_G0, 0] = rng.uniform(1e4, 3e4, 2 * nl)
    ls_X[:, abi.VAR_G0, 1] = rng.uniform(0.5, 1.0, 2 * nl)
    ls_type[:, abi.VAR_D0] = abi.TM_T1
    ls_X[:, abi.VAR_D0, 0] = rng.uniform(-500.0, 500.0, 2 * nl)
    ls_X[:, abi.VAR_D0, 1] = rng.uniform(0.5, 1.0, 2 * nl)
    return HostCatalog(
        n_species=n_species, isot_species=np.arange(n_species), isot_mass=masses, band_isot=band_isot,
        band_offset=band_offset, f0=f0, a=a, e0=e0, gu=np.full(nl, 3.0), gl=np.full(nl, 1.0),
        T0=np.full(nl, 296.0), ls_offset=ls_offset, ls_species=ls_species, ls_type=ls_type, ls_X=ls_X,
        band_cutoff_type=None if cutoff is None else np.full(nb, abi.CUTOFF_BYLINE),
        band_cutoff_value=None if cutoff is None else np.full(nb, cutoff),
    )


def _nadir_atmosphere(np_, n_species, vmr0, q0=215.0, z_top_km=None):
    """Level 0 is the sensor end (top of the atmosphere), level np-1 the surface (nadir view)."""
    z_top_km = float(np_ - 1) if z_top_km is None else z_top_km
    z = np.linspace(z_top_km, 0.0, np_)
    T, P = standard_profile(z)
    vmr = np.tile(np.asarray(vmr0, float), (np_, 1))
    vmr[:, 0] *= np.exp(-z / 2.0)  # H2O-like scale height for species 0
    isorat = np.full((np_, n_species), 0.995)
    Q = (q0 * T / 296.0)[:, None] * np.ones((1, n_species))
    dQdT = np.full((np_, n_species), q0 / 296.0)
    los = np.tile(np.array([180.0, 0.0]), (np_, 1))
    atm = AtmPath(T=T, P=P, vmr=vmr, isorat=isorat, Q=Q, dQdT=dQdT, los=los)
    r = np.abs(np.diff(z)) * 1e3
    return atm, r


def case_c2(lines_per_species=20_000, nf=100_000, np_=100, seed=2, rte_option="linsrc", bands_per_species=20,
            targets=()) -> Case:
    """C2: 100-level nadir clear-sky Tb, 5 species, 1e5 lines, 1e5 frequencies."""
    rng = np.random.default_rng(seed)
    cat = _multi_species_catalog(rng, 5, lines_per_species, bands_per_species, 1e9, 1000e9, _C2_MASS, _C2_STRENGTH)
    atm, r = _nadir_atmosphere(np_, 5, _C2_VMR0)
    f = np.linspace(1e9, 1000e9, nf)
    I_bkg = np.zeros((nf, 4))
    I_bkg[:, 0] = planck(f, 288.0)  # blackbody surface
    return Case("C2", cat, f, atm, r, I_bkg, rte_option=rte_option, targets=tuple(targets),
                note="5 species, scalar K")


def case_c4(n_lines=1_000_000, nf=1_000_000, np_=100, seed=4, cutoff=None, f_slice=None) -> Case:
    """C4: HITRAN-scale IR broadband, 1 species x 1e6 lines, 1e6 frequencies, 100 levels.

    ``f_slice=(lo, hi)`` keeps only that contiguous index range of the full grid (the
    frequency shard of one GPU, or a bounded CPU sample).
    """
    rng = np.random.default_rng(seed)
    cat = _multi_species_catalog(rng, 1, n_lines, 1, 1e12, 100e12, [43.99], [0.1], decades=6.0, cutoff=cutoff)
    atm, r = _nadir_atmosphere(np_, 1, [4e-4])
    f = np.linspace(1e12, 100e12, nf)
    if f_slice is not None:
        f = f[f_slice[0]:f_slice[1]].copy()
    I_bkg = np.zeros((len(f), 4))
    I_bkg[:, 0] = planck(f, 288.0)
    return Case("C4", cat, f, atm, r, I_bkg, note="IR broadband")


# ---------------------------------------------------------------------------
def o2_zeeman_catalog(seed=3, with_mixing=True):
    """O2-66 60 GHz-band-like Zeeman catalog: N = 1,3,..,37, N+ and N- branches.

    Upper level J = N, lower level J = N+1 (N+) or N-1 (N-); Lande factors from the
    Hund case (b) formula the reference uses (lbl_zeeman.h:178-197) with Lambda = 0, S = 1.
    """
    rng = np.random.default_rng(seed)
    GS = 2.002064
    rows = []
    for N in range(1, 38, 2):
        for branch in (+1, -1):
            Ju, Jl = N, N + branch
            if branch > 0:
                f0 = 56.2648e9 + 12.0e9 * (1 - np.exp(-(N - 1) / 14.0))
            else:
                f0 = 118.750348e9 if N == 1 else 62.4863e9 - 13.0e9 * (1 - np.exp(-(N - 3) / 16.0))
            g_of = lambda J: 0.0 if J == 0 else (GS / (N * (N + 1)) if J == N else (GS / (N + 1) if J == N + 1 else -GS / N))
            rows.append(dict(f0=f0, Ju=Ju, Jl=Jl, gu_z=g_of(Ju), gl_z=g_of(Jl), N=N))
    rows.sort(key=lambda r: r["f0"])
    nl = len(rows)
    f0 = np.array([r["f0"] for r in rows])
    N = np.array([r["N"] for r in rows], float)
    a = 4.479289583303983e-09 * (f0 / 118.750348e9) ** 3 * (1 + 0.2 * rng.uniform(0, 1, nl))
    e0 = 2.856e-23 * N * (N + 1)
    Ju = np.array([r["Ju"] for r in rows])
    Jl = np.array([r["Jl"] for r in rows])
    ls_offset, ls_species, ls_type, ls_X = _ls_tables(
        nl, rng, 0, g0=(1.2e4, 1.6e4), g0_x1=(0.75, 0.85), y=(-4e-6, 4e-6) if with_mixing else None)
    cat = HostCatalog(
        n_species=1, isot_species=[0], isot_mass=[31.9898], band_isot=[0], band_offset=[0, nl],
        f0=f0, a=a, e0=e0, gu=2.0 * Ju + 1, gl=2.0 * Jl + 1, T0=np.full(nl, 296.0),
        ls_offset=ls_offset, ls_species=ls_species, ls_type=ls_type, ls_X=ls_X,
        z_on=np.ones(nl, np.uint8), z_gu=[r["gu_z"] for r in rows], z_gl=[r["gl_z"] for r in rows],
        two_Ju=2 * Ju, two_Jl=2 * Jl,
    )
    return cat


def case_c3(nf=100_000, np_=50, seed=3, los=(180.0, 0.0), rte_option="linsrc", with_mixing=True) -> Case:
    """C3: O2 60 GHz Zeeman-split full 4x4 polarised propmat, 50 levels x 1e5 freqs."""
    cat = o2_zeeman_catalog(seed, with_mixing)
    nl = cat.n_lines
    per = max(2, nf // nl)
    win = np.linspace(-5e6, 5e6, per)
    f = np.sort((cat.f0[:, None] + win[None, :]).ravel())
    if len(f) < nf:  # pad with a few points above the last window to reach the requested size
        f = np.concatenate([f, f[-1] + 1e5 * np.arange(1, nf - len(f) + 1)])
    z = np.linspace(30.0, 80.0, np_)  # level 0 = sensor at 30 km, last level = 80 km against cold space
    T, P = standard_profile(z)
    vmr = np.full((np_, 1), 0.21)
    isorat = np.full((np_, 1), 0.995)
    Q = (215.0 * T / 296.0)[:, None]
    dQdT = np.full((np_, 1), 215.0 / 296.0)
    mag = np.array([10e-6, 50e-6, 1e-6])[None, :] * (1 + 0.01 * z)[:, None]
    losv = np.tile(np.asarray(los, float), (np_, 1))
    atm = AtmPath(T=T, P=P, vmr=vmr, isorat=isorat, Q=Q, dQdT=dQdT, mag=mag, los=losv)
    r = np.abs(np.diff(z)) * 1e3 / max(abs(np.cos(np.deg2rad(los[0]))), 1e-3)
    I_bkg = np.zeros((len(f), 4))
    I_bkg[:, 0] = planck(f, T_CMB)  # space background (m_background.cc:65)
    return Case("C3", cat, f, atm, r, I_bkg, rte_option=rte_option, note="Zeeman, polarised")


def case_c5_single(n_lines=10_000, nf=10_000, np_=100, seed=5, targets=(("T",), ("VMR", 0))) -> Case:
    """C5 (one path of the batch): C2 catalog thinned to 1e4 lines, T and VMR Jacobians."""
    c = case_c2(lines_per_species=n_lines // 5, nf=nf, np_=np_, seed=seed, bands_per_species=4, targets=targets)
    c.name = "C5"
    return c


def tiny_case(nl=64, nf=257, np_=6, seed=11, zeeman=False, cutoff=None, rte_option="linsrc", targets=()) -> Case:
    """Small ragged case for smoke / unit tests (sizes deliberately not multiples of any tile)."""
    rng = np.random.default_rng(seed)
    if zeeman:
        cat = o2_zeeman_catalog(seed)
        c3 = case_c3(nf=nf, np_=np_, seed=seed, rte_option=rte_option)
        c3.name = "tiny-zeeman"
        c3.targets = tuple(targets)
        return c3
    per_species = max(nl // 2, 1)
    cat = _multi_species_catalog(rng, 2, per_species, 2 if per_species % 2 == 0 else 1, 100e9, 130e9, [31.99, 18.01],
                                 [1.0, 30.0], decades=1.0, cutoff=cutoff)
    atm, r = _nadir_atmosphere(np_, 2, [1e-2, 0.21], z_top_km=40.0)
    f = np.linspace(100e9, 130e9, nf)
    I_bkg = np.zeros((nf, 4))
    I_bkg[:, 0] = planck(f, 288.0)
    return Case("tiny", cat, f, atm, r, I_bkg, rte_option=rte_option, targets=tuple(targets))


def mtckd_table(seed=5, n=2003, v0=-20.0, dv=10.0):
    """A synthetic MT_CKD-4.x-like water table (the real coefficients are external catalog data): regular wavenumbers from -20 cm-1,
    smooth band structure over six decades, temperature exponents between 0 and 9."""
    rng = np.random.default_rng(seed)
    wn = v0 + dv * np.arange(n)
    bands = sum(np.exp(-0.5 * ((wn - c) / w) ** 2) for c, w in ((0, 150), (1600, 120), (3700, 200), (5300, 150), (7200, 200), (10600, 250)))
    return dict(ref_temp=296.0, ref_press=1013.0, wavenumbers=wn, self_absco_ref=1e-20 * (1e-6 + bands) * rng.uniform(0.8, 1.2, n),
                for_absco_ref=3e-23 * (1e-6 + bands) * rng.uniform(0.8, 1.2, n), self_texp=rng.uniform(0.0, 9.0, n))
