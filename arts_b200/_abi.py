"""ctypes mirror of include/arts_b200.h (the C ABI) and the flat host containers.

The containers are the flattened forms a reference-side shim produces from
``AbsorptionBands`` (src/core/lbl/lbl_data.h:31-68,196-300), ``ArrayOfAtmPoint``
(src/core/atm/atm_field.h:67-82) and ``ArrayOfPropagationPathPoint``
(src/core/path/path_point.h:14-35); see INTEGRATION.md for the C++ side.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

# ---- constants mirrored from the header ------------------------------------
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM = 0, 1, 2, 3, 4
SPECIES_BATH = -1
VAR_G0, VAR_D0, VAR_DV, VAR_Y, VAR_G, NVAR = 0, 1, 2, 3, 4, 5
TM_ABSENT, TM_T0, TM_T1, TM_T2, TM_T3, TM_T4, TM_T5, TM_AER, TM_DPL, TM_POLY = -1, 0, 1, 2, 3, 4, 5, 6, 7, 8
LINESHAPE_VP_LTE, LINESHAPE_OTHER, LINESHAPE_VP_LTE_MIRROR = 0, 1, 2
CUTOFF_NONE, CUTOFF_BYLINE = 0, 1
RTE_CONSTANT, RTE_LINSRC, RTE_LINPROP = 0, 1, 2
TARGET_T, TARGET_VMR = 0, 1
TARGET_WIND_U, TARGET_WIND_V, TARGET_WIND_W = 2, 3, 4
TARGET_MAG_U, TARGET_MAG_V, TARGET_MAG_W = 5, 6, 7
TARGET_LINE_F0, TARGET_LINE_E0, TARGET_LINE_A, TARGET_LINE_LS = 8, 9, 10, 11
TARGET_P = 12
TARGET_ISORAT = 13
FLAG_K_ZERO_INIT, FLAG_TRAN_EXACT, FLAG_RETURN_K, FLAG_NO_EMISSION, FLAG_WIND_ROWS_DF = 1, 2, 4, 8, 16

RTE_OPTIONS = {"constant": RTE_CONSTANT, "linsrc": RTE_LINSRC, "lintau": RTE_LINSRC, "linprop": RTE_LINPROP}

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


class CatalogDesc(C.Structure):
    _fields_ = [
        ("n_species", C.c_int32),
        ("n_isot", C.c_int32),
        ("n_bands", C.c_int32),
        ("n_lines", C.c_int64),
        ("n_ls", C.c_int64),
        ("isot_species", _i32p),
        ("isot_mass", _dp),
        ("band_isot", _i32p),
        ("band_lineshape", _i32p),
        ("band_cutoff_type", _i32p),
        ("band_cutoff_value", _dp),
        ("band_offset", _i64p),
        ("f0", _dp),
        ("a", _dp),
        ("e0", _dp),
        ("gu", _dp),
        ("gl", _dp),
        ("T0", _dp),
        ("z_on", _u8p),
        ("z_gu", _dp),
        ("z_gl", _dp),
        ("two_Ju", _i32p),
        ("two_Jl", _i32p),
        ("ls_offset", _i64p),
        ("ls_species", _i32p),
        ("ls_type", _i32p),
        ("ls_X", _dp),
    ]


class AtmPathDesc(C.Structure):
    _fields_ = [
        ("np", C.c_int32),
        ("T", _dp),
        ("P", _dp),
        ("vmr", _dp),
        ("isorat", _dp),
        ("Q", _dp),
        ("dQdT", _dp),
        ("mag", _dp),
        ("los", _dp),
        ("wind", _dp),
    ]


class Target(C.Structure):
    _fields_ = [("kind", C.c_int32), ("species", C.c_int32), ("line", C.c_int64), ("ls_var", C.c_int32), ("coeff", C.c_int32)]


class XmlIsotopologue(C.Structure):
    """ab200_xml_isotopologue: SpeciesIsotope tag -> species index and mass."""

    _fields_ = [("name", C.c_char_p), ("species", C.c_int32), ("mass", C.c_double)]


class XmlSpecies(C.Structure):
    """ab200_xml_species: broadener name as written in the file -> species index (or SPECIES_BATH)."""

    _fields_ = [("name", C.c_char_p), ("species", C.c_int32)]


class HitranIsotopologue(C.Structure):
    """ab200_hitran_isotopologue: (HITRAN molecule number, isotopologue character) -> species index and mass."""

    _fields_ = [("M", C.c_int32), ("I", C.c_char), ("species", C.c_int32), ("mass", C.c_double), ("hitran_ratio", C.c_double),
                ("Q296", C.c_double)]


class PredefSpecies(C.Structure):
    """ab200_predef_species: indices of O2, N2, H2O, CO2, liquidcloud in the VMR vector (-1: absent)."""

    _fields_ = [("o2", C.c_int32), ("n2", C.c_int32), ("h2o", C.c_int32), ("co2", C.c_int32), ("liquidcloud", C.c_int32)]


PREDEF_MODELS = {"O2-SelfContStandardType": 0, "N2-SelfContStandardType": 1, "H2O-ForeignContStandardType": 2,
                 "H2O-SelfContStandardType": 3, "H2O-PWR98": 4, "O2-PWR98": 5, "H2O-MPM89": 6, "O2-MPM89": 7, "N2-SelfContMPM93": 8,
                 "H2O-PWR2021": 9, "H2O-PWR2022": 10, "O2-PWR2021": 11, "O2-PWR2022": 12, "N2-SelfContPWR2021": 13, "O2-TRE05": 14, "O2-MPM2020": 15,
                 "liquidcloud-ELL07": 16, "H2O-ForeignContCKDMT400": 17, "H2O-SelfContCKDMT400": 18, "H2O-ForeignContCKDMT430": 19,
                 "H2O-SelfContCKDMT430": 20}


class MtckdWater(C.Structure):
    """ab200_mtckd_water: MT_CKD400::WaterData / MT_CKD430::WaterData (src/core/predefined/predef_data.h:14-42)."""

    _fields_ = [("n", C.c_int32), ("ref_temp", C.c_double), ("ref_press", C.c_double), ("wavenumbers", _dp), ("self_absco_ref", _dp),
                ("for_absco_ref", _dp), ("self_texp", _dp)]


def predef_args(models, species):
    """(int32 array of model ids, PredefSpecies) from tag names and a dict like {"O2": 1, "H2O": 0}."""
    ids = np.array([PREDEF_MODELS[m] if isinstance(m, str) else int(m) for m in models], dtype=np.int32)
    sp = PredefSpecies(*(int(species.get(k, -1)) for k in ("O2", "N2", "H2O", "CO2", "liquidcloud")))
    return ids, sp


class LookupTableDesc(C.Structure):
    _fields_ = [("species", C.c_int32), ("nf", C.c_int32), ("np", C.c_int32), ("nt", C.c_int32), ("nw", C.c_int32),
                ("do_t", C.c_int32), ("do_w", C.c_int32), ("f_grid", _dp), ("log_p_grid", _dp), ("t_pert", _dp), ("w_pert", _dp),
                ("t_atmref", _dp), ("water_atmref", _dp), ("xsec", _dp)]


@dataclass
class LookupTable:
    """lookup::table (src/core/lookup/lookup_map.h) of one species: ``xsec`` [nt, nw, np, nf]."""

    species: int
    f_grid: np.ndarray
    log_p_grid: np.ndarray
    t_atmref: np.ndarray
    xsec: np.ndarray
    t_pert: np.ndarray | None = None
    w_pert: np.ndarray | None = None
    water_atmref: np.ndarray | None = None

    def desc(self):
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        self._k = [c(self.f_grid), c(self.log_p_grid), c(self.t_pert), c(self.w_pert), c(self.t_atmref), c(self.water_atmref),
                   c(self.xsec)]
        d = LookupTableDesc()
        d.species, d.nf, d.np = int(self.species), len(self._k[0]), len(self._k[1])
        d.do_t, d.do_w = int(self.t_pert is not None), int(self.w_pert is not None)
        d.nt = len(self._k[2]) if d.do_t else 1
        d.nw = len(self._k[3]) if d.do_w else 1
        assert self._k[6].shape == (d.nt, d.nw, d.np, d.nf), (self._k[6].shape, (d.nt, d.nw, d.np, d.nf))
        d.f_grid, d.log_p_grid, d.t_pert, d.w_pert, d.t_atmref, d.water_atmref, d.xsec = (dptr(a) for a in self._k)
        return d


def lookup_tables(tables):
    arr = (LookupTableDesc * len(tables))()
    for k, t in enumerate(tables):
        arr[k] = t.desc()
    return arr


class PartfunTable(C.Structure):
    """ab200_partfun_table: kind 0 interp, 1 coeff, 2 const, 3 static_interp (src/partfun/make_auto_partfuns.cc)."""

    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("grid", _dp), ("coef", _dp)]


class AtmProfileDesc(C.Structure):
    """ab200_atm_profile: the AtmField of a 1-D atmosphere, flattened."""

    _fields_ = [("nalt", C.c_int32), ("alt", _dp), ("T", _dp), ("P", _dp), ("vmr", _dp), ("isorat", _dp), ("mag", _dp), ("wind", _dp),
                ("alt_low", C.c_int32), ("alt_upp", C.c_int32), ("top_of_atmosphere", C.c_double), ("partfun", C.POINTER(PartfunTable))]


EXTRAP = {"None": 0, "Zero": 1, "Nearest": 2, "Linear": 3}
PARTFUN_KINDS = {"interp": 0, "coeff": 1, "const": 2, "static_interp": 3}


class CiaDatasetDesc(C.Structure):
    _fields_ = [("nf", C.c_int32), ("nT", C.c_int32), ("f_grid", _dp), ("T_grid", _dp), ("data", _dp)]


class CiaRecordDesc(C.Structure):
    _fields_ = [("species1", C.c_int32), ("species2", C.c_int32), ("n_datasets", C.c_int32),
                ("datasets", C.POINTER(CiaDatasetDesc))]


@dataclass
class CiaRecord:
    """CIARecord (src/core/absorption/cia.h): one species pair with its data sets, each a GriddedField2
    ``(f_grid [nf], T_grid [nT], data [nf, nT])``."""

    species1: int
    species2: int
    datasets: list

    def desc(self):
        self._keep = [(np.ascontiguousarray(f, np.float64), np.ascontiguousarray(T, np.float64), np.ascontiguousarray(d, np.float64))
                      for (f, T, d) in self.datasets]
        self._ds = (CiaDatasetDesc * len(self._keep))()
        for k, (f, T, d) in enumerate(self._keep):
            assert d.shape == (len(f), len(T))
            self._ds[k].nf, self._ds[k].nT = len(f), len(T)
            self._ds[k].f_grid, self._ds[k].T_grid, self._ds[k].data = dptr(f), dptr(T), dptr(d)
        r = CiaRecordDesc()
        r.species1, r.species2, r.n_datasets, r.datasets = int(self.species1), int(self.species2), len(self._keep), self._ds
        return r


def cia_records(records):
    arr = (CiaRecordDesc * len(records))()
    for k, r in enumerate(records):
        arr[k] = r.desc()
    return arr


class ObserverDesc(C.Structure):
    """ab200_observer (include/arts_b200.h)."""

    _fields_ = [
        ("bkg_kind", C.c_int32),
        ("bkg_T", C.c_double),
        ("nx", C.c_int32),
        ("map_offset", C.POINTER(C.c_int64)),
        ("map_x", C.POINTER(C.c_int32)),
        ("map_w", _dp),
        ("n_bkg", C.c_int32),
        ("bkg_x", C.POINTER(C.c_int32)),
        ("bkg_w", _dp),
        ("unit", C.c_int32),
        ("n_real", C.c_double),
        ("n_channels", C.c_int32),
        ("w_offset", C.POINTER(C.c_int64)),
        ("w_freq", C.POINTER(C.c_int64)),
        ("w_stokes", _dp),
    ]


UNITS = {"unit": 0, "RJBT": 1, "PlanckBT": 2, "W_m2_m_sr": 3, "W_m2_m1_sr": 4}
BKG_UPLOADED, BKG_PLANCK = 0, 1


def _arr(x, dtype, shape=None):
    a = np.ascontiguousarray(x, dtype=dtype)
    if shape is not None:
        a = a.reshape(shape)
    return a


def ptr(a: np.ndarray | None, ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))


def dptr(a):
    return ptr(a, C.c_double)


@dataclass
class HostCatalog:
    """AbsorptionBands flattened to SoA (what ab200_catalog_create consumes).

    Lines are stored band-contiguous and, inside a band, sorted by ``f0`` like
    the reference requires for ``band_data::active_lines`` (lbl_data.cpp:61-68).
    """

    n_species: int
    isot_species: np.ndarray
    isot_mass: np.ndarray
    band_isot: np.ndarray
    band_offset: np.ndarray
    f0: np.ndarray
    a: np.ndarray
    e0: np.ndarray
    gu: np.ndarray
    gl: np.ndarray
    T0: np.ndarray
    ls_offset: np.ndarray
    ls_species: np.ndarray
    ls_type: np.ndarray  # [n_ls, NVAR]
    ls_X: np.ndarray  # [n_ls, NVAR, 4]
    band_lineshape: np.ndarray | None = None
    band_cutoff_type: np.ndarray | None = None
    band_cutoff_value: np.ndarray | None = None
    z_on: np.ndarray | None = None
    z_gu: np.ndarray | None = None
    z_gl: np.ndarray | None = None
    two_Ju: np.ndarray | None = None
    two_Jl: np.ndarray | None = None
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        nb = len(self.band_isot)
        nl = len(self.f0)
        self.isot_species = _arr(self.isot_species, np.int32)
        self.isot_mass = _arr(self.isot_mass, np.float64)
        self.band_isot = _arr(self.band_isot, np.int32)
        self.band_offset = _arr(self.band_offset, np.int64)
        for name in ("f0", "a", "e0", "gu", "gl", "T0"):
            setattr(self, name, _arr(getattr(self, name), np.float64))
        self.ls_offset = _arr(self.ls_offset, np.int64)
        self.ls_species = _arr(self.ls_species, np.int32)
        nls = len(self.ls_species)
        self.ls_type = _arr(self.ls_type, np.int32, (nls, NVAR))
        self.ls_X = _arr(self.ls_X, np.float64, (nls, NVAR, 4))
        self.band_lineshape = _arr(np.zeros(nb) if self.band_lineshape is None else self.band_lineshape, np.int32)
        self.band_cutoff_type = _arr(np.zeros(nb) if self.band_cutoff_type is None else self.band_cutoff_type, np.int32)
        self.band_cutoff_value = _arr(
            np.full(nb, np.inf) if self.band_cutoff_value is None else self.band_cutoff_value, np.float64
        )
        self.z_on = _arr(np.zeros(nl) if self.z_on is None else self.z_on, np.uint8)
        self.z_gu = _arr(np.zeros(nl) if self.z_gu is None else self.z_gu, np.float64)
        self.z_gl = _arr(np.zeros(nl) if self.z_gl is None else self.z_gl, np.float64)
        self.two_Ju = _arr(np.zeros(nl) if self.two_Ju is None else self.two_Ju, np.int32)
        self.two_Jl = _arr(np.zeros(nl) if self.two_Jl is None else self.two_Jl, np.int32)
        assert self.band_offset.shape == (nb + 1,) and self.band_offset[-1] == nl
        assert self.ls_offset.shape == (nl + 1,) and self.ls_offset[-1] == nls

    @property
    def n_lines(self) -> int:
        return int(len(self.f0))

    @property
    def n_isot(self) -> int:
        return int(len(self.isot_species))

    @property
    def n_bands(self) -> int:
        return int(len(self.band_isot))

    def desc(self) -> CatalogDesc:
        d = CatalogDesc()
        d.n_species = self.n_species
        d.n_isot = self.n_isot
        d.n_bands = self.n_bands
        d.n_lines = self.n_lines
        d.n_ls = len(self.ls_species)
        d.isot_species = ptr(self.isot_species, C.c_int32)
        d.isot_mass = dptr(self.isot_mass)
        d.band_isot = ptr(self.band_isot, C.c_int32)
        d.band_lineshape = ptr(self.band_lineshape, C.c_int32)
        d.band_cutoff_type = ptr(self.band_cutoff_type, C.c_int32)
        d.band_cutoff_value = dptr(self.band_cutoff_value)
        d.band_offset = ptr(self.band_offset, C.c_int64)
        for name in ("f0", "a", "e0", "gu", "gl", "T0", "z_gu", "z_gl"):
            setattr(d, name, dptr(getattr(self, name)))
        d.z_on = ptr(self.z_on, C.c_uint8)
        d.two_Ju = ptr(self.two_Ju, C.c_int32)
        d.two_Jl = ptr(self.two_Jl, C.c_int32)
        d.ls_offset = ptr(self.ls_offset, C.c_int64)
        d.ls_species = ptr(self.ls_species, C.c_int32)
        d.ls_type = ptr(self.ls_type, C.c_int32)
        d.ls_X = dptr(self.ls_X)
        return d


@dataclass
class AtmPath:
    """ArrayOfAtmPoint + the los of ArrayOfPropagationPathPoint, flattened per level."""

    T: np.ndarray
    P: np.ndarray
    vmr: np.ndarray  # [np, n_species]
    isorat: np.ndarray  # [np, n_isot]
    Q: np.ndarray  # [np, n_isot]
    dQdT: np.ndarray | None = None
    mag: np.ndarray | None = None  # [np, 3]
    los: np.ndarray | None = None  # [np, 2]
    wind: np.ndarray | None = None  # [np, 3] u, v, w [m/s]

    def __post_init__(self):
        self.T = _arr(self.T, np.float64)
        n = len(self.T)
        self.P = _arr(self.P, np.float64)
        self.vmr = _arr(self.vmr, np.float64).reshape(n, -1)
        self.isorat = _arr(self.isorat, np.float64).reshape(n, -1)
        self.Q = _arr(self.Q, np.float64).reshape(n, -1)
        if self.dQdT is not None:
            self.dQdT = _arr(self.dQdT, np.float64).reshape(n, -1)
        if self.mag is not None:
            self.mag = _arr(self.mag, np.float64).reshape(n, 3)
        if self.los is not None:
            self.los = _arr(self.los, np.float64).reshape(n, 2)
        if self.wind is not None:
            self.wind = _arr(self.wind, np.float64).reshape(n, 3)

    @property
    def np_(self) -> int:
        return int(len(self.T))

    def reversed(self) -> "AtmPath":
        """The same points in the opposite order (a path runs sensor -> background, a lookup-table profile surface -> top)."""
        r = lambda a: None if a is None else np.ascontiguousarray(a[::-1])  # noqa: E731
        return AtmPath(T=r(self.T), P=r(self.P), vmr=r(self.vmr), isorat=r(self.isorat), Q=r(self.Q), dQdT=r(self.dQdT), mag=r(self.mag),
                       los=r(self.los), wind=r(self.wind))

    def take(self, idx) -> "AtmPath":
        """The points ``idx`` (any numpy index) as a new path."""
        r = lambda a: None if a is None else np.ascontiguousarray(a[idx])  # noqa: E731
        return AtmPath(T=r(self.T), P=r(self.P), vmr=r(self.vmr), isorat=r(self.isorat), Q=r(self.Q), dQdT=r(self.dQdT), mag=r(self.mag),
                       los=r(self.los), wind=r(self.wind))

    def level(self, ip: int) -> "AtmPath":
        s = slice(ip, ip + 1)
        return AtmPath(
            self.T[s], self.P[s], self.vmr[s], self.isorat[s], self.Q[s],
            None if self.dQdT is None else self.dQdT[s],
            None if self.mag is None else self.mag[s],
            None if self.los is None else self.los[s],
            None if self.wind is None else self.wind[s],
        )

    def desc(self) -> AtmPathDesc:
        d = AtmPathDesc()
        d.np = self.np_
        d.T = dptr(self.T)
        d.P = dptr(self.P)
        d.vmr = dptr(self.vmr)
        d.isorat = dptr(self.isorat)
        d.Q = dptr(self.Q)
        d.dQdT = dptr(self.dQdT)
        d.mag = dptr(self.mag)
        d.los = dptr(self.los)
        d.wind = dptr(self.wind)
        return d


@dataclass
class Observer:
    """The host glue around one path, flattened (ab200_observer): background, the path-point -> state-vector map of
    ``spectral_rad_jacAddPathPropagation`` (src/m_rad.cc:62-127), the unit of ``spectral_rad_transform_operator`` and
    this path's rows of the sensor's sparse weight matrices (``SensorObsel::sumup``, src/core/sensor/obsel.cpp:246-279).

    ``path_map[ip][t]`` is a list of ``(x index, weight)``; ``bkg_rows`` a list of ``(x index, weight)`` for the
    surface-temperature target; ``channels[c]`` a list of ``(frequency index, (wI, wQ, wU, wV))``.
    """

    nx: int
    path_map: list | None = None
    bkg_T: float | None = None  # None: the uploaded I_bkg
    bkg_rows: list = field(default_factory=list)
    unit: str = "unit"
    n_real: float = 1.0
    channels: list = field(default_factory=list)

    def desc(self, np_: int, nq: int) -> ObserverDesc:
        cached = getattr(self, "_desc_cache", None)
        if cached is not None and cached[0] == (np_, nq, self.nx, self.bkg_T, self.unit, self.n_real):
            return cached[1]  # the flattening of the Python lists is the expensive part; observers are reused per path
        d = self._build_desc(np_, nq)
        self._desc_cache = ((np_, nq, self.nx, self.bkg_T, self.unit, self.n_real), d)
        return d

    def _build_desc(self, np_: int, nq: int) -> ObserverDesc:
        off, xs, ws = [0], [], []
        for ip in range(np_):
            for t in range(nq):
                for (x, w) in (self.path_map[ip][t] if self.path_map else ()):
                    xs.append(int(x)); ws.append(float(w))
                off.append(len(xs))
        self._map_offset = np.asarray(off, np.int64)
        self._map_x = np.asarray(xs, np.int32)
        self._map_w = np.asarray(ws, np.float64)
        self._bkg_x = np.asarray([r[0] for r in self.bkg_rows], np.int32)
        self._bkg_w = np.asarray([r[1] for r in self.bkg_rows], np.float64)
        woff, wf, wst = [0], [], []
        for ch in self.channels:
            for (j, w4) in ch:
                wf.append(int(j)); wst.append([float(v) for v in w4])
            woff.append(len(wf))
        self._w_offset = np.asarray(woff, np.int64)
        self._w_freq = np.asarray(wf, np.int64)
        self._w_stokes = np.asarray(wst, np.float64).reshape(-1, 4)
        d = ObserverDesc()
        d.bkg_kind = BKG_UPLOADED if self.bkg_T is None else BKG_PLANCK
        d.bkg_T = 0.0 if self.bkg_T is None else float(self.bkg_T)
        d.nx = int(self.nx)
        d.map_offset = ptr(self._map_offset, C.c_int64)
        d.map_x = ptr(self._map_x, C.c_int32)
        d.map_w = dptr(self._map_w)
        d.n_bkg = len(self._bkg_x)
        d.bkg_x = ptr(self._bkg_x, C.c_int32)
        d.bkg_w = dptr(self._bkg_w)
        d.unit = UNITS[self.unit]
        d.n_real = float(self.n_real)
        d.n_channels = len(self.channels)
        d.w_offset = ptr(self._w_offset, C.c_int64)
        d.w_freq = ptr(self._w_freq, C.c_int64)
        d.w_stokes = dptr(self._w_stokes)
        return d


def make_targets(targets) -> tuple[C.Array | None, int]:
    """targets: iterable of ("T",) / ("VMR", species_id) / ("wind_u",) .. ("wind_w",) / ("mag_u",) .. ("mag_w",) /
    ("line_f0", line) / ("line_e0", line) / ("line_a", line) / ("line_ls", line, var, species, coeff) or (kind, species) ints."""
    lst = []
    for t in targets or ():
        if isinstance(t, str):
            t = (t,)
        kind = t[0]
        if kind in ("T", "t"):
            lst.append((TARGET_T, 0, 0, 0, 0))
        elif kind in ("VMR", "vmr"):
            lst.append((TARGET_VMR, int(t[1]), 0, 0, 0))
        elif kind in ("wind_u", "wind_v", "wind_w"):  # AtmKey::wind_*
            lst.append((TARGET_WIND_U + "uvw".index(kind[-1]), 0, 0, 0, 0))
        elif kind in ("mag_u", "mag_v", "mag_w"):  # AtmKey::mag_*
            lst.append((TARGET_MAG_U + "uvw".index(kind[-1]), 0, 0, 0, 0))
        elif kind in ("line_f0", "line_e0", "line_a"):  # lbl::line_key with LineByLineVariable
            lst.append(({"line_f0": TARGET_LINE_F0, "line_e0": TARGET_LINE_E0, "line_a": TARGET_LINE_A}[kind], 0, int(t[1]), 0, 0))
        elif kind in ("isorat", "ISORAT"):  # SpeciesIsotope ratio: ("isorat", isotopologue index)
            lst.append((TARGET_ISORAT, int(t[1]), 0, 0, 0))
        elif kind in ("p", "P"):  # AtmKey::p: the reference's "Not implemented, pressure derivative"
            lst.append((TARGET_P, 0, 0, 0, 0))
        elif kind == "line_ls":  # lbl::line_key with LineShapeModelVariable, broadener and coefficient
            lst.append((TARGET_LINE_LS, int(t[3]), int(t[1]), int(t[2]), int(t[4])))
        else:
            lst.append((int(kind), int(t[1]) if len(t) > 1 else 0, 0, 0, 0))
    if not lst:
        return None, 0
    arr = (Target * len(lst))()
    for i, (k, s, ln, var, co) in enumerate(lst):
        arr[i].kind = k
        arr[i].species = s
        arr[i].line = ln
        arr[i].ls_var = var
        arr[i].coeff = co
    return arr, len(lst)


# argument lists shared by the CUDA library (ab200_*) and the oracle (orc_*)
SIG_PROPMAT_LEVELS_CORE = [
    C.c_int64, _dp, C.c_int64, C.POINTER(AtmPathDesc), C.c_int32, C.c_int32, C.c_int32, C.POINTER(Target),
]
SIG_TRAMAT = [C.c_int32, C.c_int64, C.c_int32, _dp, _dp, _dp, _dp, C.c_int32, C.c_uint32, _dp, _dp, _dp, _dp, _dp]
SIG_SRCVEC = [C.c_int32, C.c_int64, C.c_int32, _dp, _dp, C.c_int64, _dp, C.c_int32, _dp, _dp]
SIG_RTE = [C.c_int32, C.c_int32, C.c_int64, C.c_int32, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
SIG_TRANSMISSION = [C.c_int32, C.c_int64, C.c_int32, _dp, _dp, _dp, _dp, _dp, _dp]
SIG_CLEARSKY_CORE = [
    C.c_int64, _dp, C.c_int64, C.POINTER(AtmPathDesc), C.c_int32, C.c_int32, C.c_int32, C.POINTER(Target), _dp,
    C.c_int32, C.c_int32, _dp, C.c_uint32, _dp, _dp, _dp,
]
