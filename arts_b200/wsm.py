"""Host-side mirror of the reference's workspace methods on the clear-sky spectral path.

Same names, argument meaning and error behaviour as the ARTS WSMs, on the flattened
containers of ``arts_b200._abi`` and numpy arrays with the reference's storage layouts:

=====================================  ===========================================
reference WSM (file:line)              here
=====================================  ===========================================
spectral_propmatAddLines               :func:`spectral_propmatAddLines`
  (src/m_lbl.cc:242-300)
spectral_propmat_pathFromPath          :func:`spectral_propmat_pathFromPath`
  (src/m_propmat.cc:5-65)
spectral_tramat_pathFromPath           :func:`spectral_tramat_pathFromPath`
  (src/m_tramat.cc:3-27)
spectral_rad_srcvec_pathFromPropmat    :func:`spectral_rad_srcvec_pathFromPropmat`
  (src/m_srcvec.cc:8-31)
spectral_radStepByStepEmission         :func:`spectral_radStepByStepEmission`
  (src/m_spectral_radiance.cc:18-46)
spectral_radClearskyEmission           :func:`spectral_radClearskyEmission` (fused)
  (src/workspace_meta_methods.cpp:166-181)
spectral_radApplyUnit... (PlanckBT)    :func:`spectral_radApplyPlanckTb`
=====================================  ===========================================

Everything calls the CUDA library through the C ABI (``include/arts_b200.h``); nothing
here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi as abi
from ._abi import AtmPath, HostCatalog, dptr, make_targets
from ._lib import Ab200Error, check, lib  # noqa: F401


def device_count() -> int:
    return int(lib().ab200_device_count())


def set_device(device: int):
    check(lib().ab200_set_device(int(device)))


def set_thread_stream(stream):
    """cudaStream_t (int) for this thread's host-buffer entry points, or None for the library's own."""
    check(lib().ab200_set_thread_stream(C.c_void_p(int(stream)) if stream else C.c_void_p()))


class Catalog:
    """Device-resident ``AbsorptionBands`` (ab200_catalog); immutable, shareable."""

    def __init__(self, host: HostCatalog):
        self.host = host
        self._h = C.c_void_p()
        d = host.desc()
        check(lib().ab200_catalog_create(C.byref(d), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def counts(self):
        out = (C.c_int64 * 4)()
        check(lib().ab200_catalog_counts(self._h, out))
        return list(out)

    def close(self):
        if self._h:
            lib().ab200_release_thread_cache()
            lib().ab200_catalog_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_catalog(abs_bands):
    if isinstance(abs_bands, Catalog) or type(abs_bands).__name__ == "MultiDevice":
        return abs_bands
    return Catalog(abs_bands)


def _f_arg(f, np_):
    f = np.ascontiguousarray(f, dtype=np.float64)
    if f.ndim == 1:
        return f, 0, f.shape[0]
    if f.shape[0] != np_:
        raise ValueError(f"Not same size: freq_grid_path {f.shape[0]} element(s), atm_path {np_} element(s)")
    return f, f.shape[1], f.shape[1]


def _rte(option) -> int:
    if isinstance(option, str):
        if option not in abi.RTE_OPTIONS:
            raise ValueError(f"Bad TransmittanceOption: {option}")
        return abi.RTE_OPTIONS[option]
    return int(option)


# ---------------------------------------------------------------------------
def spectral_propmatAddLines(spectral_propmat, spectral_propmat_jac, freq_grid, jac_targets, select_species,
                             abs_bands, atm_point: AtmPath, no_negative_absorption=1):
    """``spectral_propmat[nf,7] += lines`` at one atmospheric point (src/m_lbl.cc:242-300).  ``freq_grid`` is the point's
    (already wind-shifted) grid and wind rows are left as the frequency derivative, like in the reference, whose agenda
    runs ``spectral_propmat_jacWindFix`` afterwards."""
    if atm_point.np_ != 1:
        raise ValueError("spectral_propmatAddLines takes a single AtmPoint")
    K = spectral_propmat.reshape(1, *spectral_propmat.shape)
    dK = None if spectral_propmat_jac is None else spectral_propmat_jac.reshape(1, *spectral_propmat_jac.shape)
    spectral_propmat_pathFromPath(abs_bands, freq_grid, atm_point, jac_targets=jac_targets,
                                  select_species=select_species, no_negative_absorption=no_negative_absorption,
                                  out=K, out_jac=dK, accumulate=True, wind_rows_df=True)
    return spectral_propmat


class MultiDevice:
    """One host process, several GPUs (ab200_multi): a catalog replica and a worker thread per device.  Pass it where a
    catalog goes in ``spectral_radClearskyEmission`` / ``spectral_propmat_pathFromPath``: the call's frequency grid is
    dealt over the devices in 512-frequency blocks and every device writes its blocks into the caller's arrays — the
    reference's OpenMP frequency loop (src/m_lbl.cc:273-295) with devices for threads."""

    def __init__(self, host: HostCatalog, n_devices: int = 0, devices=None):
        self.host = host
        self._h = C.c_void_p()
        d = host.desc()
        dev = None
        if devices is not None:
            n_devices = len(devices)
            dev = (C.c_int32 * n_devices)(*[int(x) for x in devices])
        check(lib().ab200_multi_create(C.byref(d), int(n_devices), dev, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    @property
    def n_devices(self):
        return int(lib().ab200_multi_device_count(self._h))

    def close(self):
        if self._h:
            lib().ab200_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spectral_propmat_pathFromPath(abs_bands, freq_grid_path, atm_path: AtmPath, jac_targets=(),
                                  select_species=abi.SPECIES_BATH, no_negative_absorption=1, out=None, out_jac=None,
                                  accumulate=False, wind_rows_df=False):
    """All levels of ``spectral_propmat_path`` with the lines-only agenda (src/m_propmat.cc:5-65), its closing
    ``spectral_propmat_jacWindFix`` included unless ``wind_rows_df``.

    Returns ``(K[np,nf,7], dK[np,nq,nf,7])``.
    """
    cat = _as_catalog(abs_bands)
    f, stride, nf = _f_arg(freq_grid_path, atm_path.np_)
    tg, nq = make_targets(jac_targets)
    np_ = atm_path.np_
    K = np.zeros((np_, nf, 7)) if out is None else out
    dK = (np.zeros((np_, nq, nf, 7)) if out_jac is None else out_jac) if nq else None
    if K.shape != (np_, nf, 7) or not K.flags.c_contiguous or K.dtype != np.float64:
        raise ValueError("spectral_propmat must be a C-contiguous float64 [np, nf, 7] array")
    flags = 0 if accumulate else abi.FLAG_K_ZERO_INIT
    if wind_rows_df:
        flags |= abi.FLAG_WIND_ROWS_DF
    a = atm_path.desc()
    fn = lib().ab200_multi_propmat_levels if isinstance(cat, MultiDevice) else lib().ab200_propmat_levels
    check(fn(cat.handle, nf, dptr(f), stride, C.byref(a), int(select_species),
             int(no_negative_absorption), nq, tg, flags, dptr(K), dptr(dK)))
    return K, dK


@dataclass
class TransmittanceMatrix:
    """rtepack::TransmittanceMatrix (rtepack_transmission.h:10-25): T, L, P [nf,np,4,4], dT, dL [2,nf,np,nq,4,4]."""

    option: str
    T: np.ndarray
    L: np.ndarray | None
    P: np.ndarray
    dT: np.ndarray
    dL: np.ndarray | None


def spectral_tramat_pathFromPath(spectral_propmat_path, spectral_propmat_jac_path, r, atm_T, rte_option="linsrc",
                                 hse_derivative=0, it=-1, flags=0) -> TransmittanceMatrix:
    """src/m_tramat.cc:3-27; ``r`` = distance(ray_path) [np-1]; ``it`` = position of the T target."""
    K = np.ascontiguousarray(spectral_propmat_path, dtype=np.float64)
    np_, nf, _ = K.shape
    dK = spectral_propmat_jac_path
    nq = 0 if dK is None else dK.shape[1]
    r = np.ascontiguousarray(r, dtype=np.float64)
    if r.shape != (max(np_ - 1, 0),):
        raise ValueError(f"dr and r must have compatible sizes. r: {r.shape}, np-1: {np_ - 1}, nq: {nq}")
    dr = np.zeros((2, max(np_ - 1, 0), nq))
    if hse_derivative and it >= 0:  # m_tramat.cc:18-24
        atm_T = np.asarray(atm_T, dtype=np.float64)
        dr[0, :, it] = r / (2.0 * atm_T[:-1])
        dr[1, :, it] = r / (2.0 * atm_T[1:])
    opt = _rte(rte_option)
    T = np.empty((nf, np_, 4, 4))
    P = np.empty((nf, np_, 4, 4))
    L = np.empty((nf, np_, 4, 4)) if opt != abi.RTE_CONSTANT else None
    dT = np.empty((2, nf, np_, nq, 4, 4))
    dL = np.empty((2, nf, np_, nq, 4, 4)) if opt != abi.RTE_CONSTANT else None
    dKc = None if nq == 0 else np.ascontiguousarray(dK, dtype=np.float64)
    if nq:
        dT[...] = 0.0  # the library fills only the entries the reference writes (rest: muelmat::constant(0))
        if dL is not None:
            dL[...] = 0.0
    check(lib().ab200_tramat(np_, nf, nq, dptr(K), dptr(dKc), dptr(r), dptr(dr), opt, flags, dptr(T), dptr(L), dptr(P),
                             dptr(dT), dptr(dL)))
    name = rte_option if isinstance(rte_option, str) else {0: "constant", 1: "linsrc", 2: "linprop"}[opt]
    return TransmittanceMatrix(name, T, L, P, dT, dL)


def spectral_rad_srcvec_pathFromPropmat(spectral_propmat_path, freq_grid_path, atm_T, it=-1, nq=0):
    """src/m_srcvec.cc:8-31 in LTE: returns ``(J[nf,np,4], dJ[nf,np,nq,4])``."""
    K = np.ascontiguousarray(spectral_propmat_path, dtype=np.float64)
    np_, nf, _ = K.shape
    f, stride, nf2 = _f_arg(freq_grid_path, np_)
    if nf2 != nf:
        raise ValueError("All forward parameters must have same shape")
    T = np.ascontiguousarray(atm_T, dtype=np.float64)
    J = np.empty((nf, np_, 4))
    dJ = np.empty((nf, np_, nq, 4))
    check(lib().ab200_srcvec(np_, nf, nq, dptr(K), dptr(f), stride, dptr(T), it, dptr(J), dptr(dJ)))
    return J, dJ


def spectral_radStepByStepEmission(spectral_tramat: TransmittanceMatrix, J, dJ, spectral_rad_bkg):
    """src/m_spectral_radiance.cc:18-46: returns ``(spectral_rad[nf,4], spectral_rad_jac_path[nf,np,nq,4])``."""
    T = spectral_tramat.T
    nf, np_ = T.shape[:2]
    nq = dJ.shape[2] if dJ is not None else 0
    bkg = np.ascontiguousarray(spectral_rad_bkg, dtype=np.float64)
    if bkg.shape[0] != nf:
        raise ValueError(f"Bad background radiance size: spectral_rad_bkg: {bkg.shape[0]}, expected: {nf}")
    I = np.empty((nf, 4))
    dI = np.zeros((nf, np_, nq, 4))
    Jc = np.ascontiguousarray(J, dtype=np.float64)
    dJc = None if dJ is None else np.ascontiguousarray(dJ, dtype=np.float64)
    check(lib().ab200_rte_emission(_rte(spectral_tramat.option), np_, nf, nq, dptr(T), dptr(spectral_tramat.L),
                                   dptr(spectral_tramat.P), dptr(spectral_tramat.dT), dptr(spectral_tramat.dL),
                                   dptr(Jc), dptr(dJc), dptr(bkg), dptr(I), dptr(dI)))
    return I, dI


def spectral_radCumulativeTransmission(spectral_tramat: TransmittanceMatrix, spectral_rad_bkg):
    """src/m_spectral_radiance.cc:49-74 (``rte_transmission``, rtepack_rtestep.cc:456-503): the background radiance
    seen through the path, ``P[:, np-1] @ spectral_rad_bkg``, and its transmission-only ``spectral_rad_jac_path``."""
    P = spectral_tramat.P
    nf, np_ = P.shape[:2]
    dT = spectral_tramat.dT
    nq = 0 if dT is None else dT.shape[3]
    bkg = np.ascontiguousarray(spectral_rad_bkg, dtype=np.float64)
    if bkg.shape[0] != nf:
        raise ValueError(f"Bad background radiance size: spectral_rad_bkg: {bkg.shape[0]}, expected: {nf}")
    I = np.empty((nf, 4))
    dI = np.zeros((nf, np_, nq, 4))
    check(lib().ab200_rte_transmission(np_, nf, nq, dptr(spectral_tramat.T), dptr(P), dptr(dT if nq else None), dptr(bkg),
                                       dptr(I), dptr(dI)))
    return I, dI


def spectral_radClearskyEmission(abs_bands, freq_grid_path, atm_path: AtmPath, r, spectral_rad_bkg,
                                 rte_option="linsrc", jac_targets=(), select_species=abi.SPECIES_BATH,
                                 no_negative_absorption=1, hse_derivative=0, flags=0, return_propmat=False):
    """The fused path: propagation matrix -> transmission -> source -> recursion on the device.

    Returns ``(spectral_rad[nf,4], spectral_rad_jac_path | None[, spectral_propmat_path])``.
    """
    cat = _as_catalog(abs_bands)
    np_ = atm_path.np_
    f, stride, nf = _f_arg(freq_grid_path, np_)
    tg, nq = make_targets(jac_targets)
    r = np.ascontiguousarray(r, dtype=np.float64)
    bkg = np.ascontiguousarray(spectral_rad_bkg, dtype=np.float64)
    if bkg.shape != (nf, 4):
        raise ValueError(f"Bad background radiance size: spectral_rad_bkg: {bkg.shape[0]}, expected: {nf}")
    if r.shape != (max(np_ - 1, 0),):
        raise ValueError(f"r must have np-1 = {np_ - 1} elements, has {r.shape}")
    I = np.empty((nf, 4))
    dI = np.empty((nf, np_, nq, 4)) if nq else None
    K = np.empty((np_, nf, 7)) if return_propmat else None
    if return_propmat:
        flags |= abi.FLAG_RETURN_K
    a = atm_path.desc()
    fn = lib().ab200_multi_clearsky_emission if isinstance(cat, MultiDevice) else lib().ab200_clearsky_emission
    check(fn(cat.handle, nf, dptr(f), stride, C.byref(a), int(select_species),
             int(no_negative_absorption), nq, tg, dptr(r), int(hse_derivative),
             _rte(rte_option), dptr(bkg), flags, dptr(I), dptr(dI), dptr(K)))
    return (I, dI, K) if return_propmat else (I, dI)


def spectral_radApplyPlanckTb(spectral_rad, freq_grid):
    """PlanckBT unit conversion (spectral_radiance_transform_operator.cc:46-87), returns a new array."""
    f = np.ascontiguousarray(freq_grid, dtype=np.float64)
    out = np.array(spectral_rad, dtype=np.float64, order="C", copy=True)
    check(lib().ab200_planck_tb(len(f), dptr(f), dptr(out)))
    return out


def faddeeva_w(z):
    """Device w(z) (tests)."""
    z = np.ascontiguousarray(z, dtype=np.complex128).ravel()
    zr, zi = np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)
    wr, wi = np.empty_like(zr), np.empty_like(zr)
    check(lib().ab200_faddeeva_w(len(zr), dptr(zr), dptr(zi), dptr(wr), dptr(wi)))
    return wr + 1j * wi


def dawson(z):
    """Device Faddeeva::Dawson(z), complex z (tests)."""
    z = np.ascontiguousarray(z, dtype=np.complex128).ravel()
    zr, zi = np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)
    dr, di = np.empty_like(zr), np.empty_like(zr)
    check(lib().ab200_dawson(len(zr), dptr(zr), dptr(zi), dptr(dr), dptr(di)))
    return dr + 1j * di


def measure_dfma_peak(iters=20000):
    t, ms = C.c_double(), C.c_double()
    check(lib().ab200_measure_dfma_peak(int(iters), C.byref(t), C.byref(ms)))
    return t.value, ms.value


def measure_dfma_mix(iters=20000):
    t, ms = C.c_double(), C.c_double()
    check(lib().ab200_measure_dfma_mix(int(iters), C.byref(t), C.byref(ms)))
    return t.value, ms.value


class Path:
    """Device-resident workspace of one propagation path (ab200_path): upload once, run, download."""

    def __init__(self, cat: Catalog, nf: int, np_: int, nq: int = 0, stream=None, stage2_only: bool = False):
        self.cat, self.nf, self.np_, self.nq = cat, int(nf), int(np_), int(nq)
        self.np_cap = int(np_)
        self.k_pitch = (self.nf + 127) // 128 * 128  # frequencies per level row of the resident K [np][k_pitch][7]
        self._h = C.c_void_p()
        if stage2_only:  # Stokes chain only; K comes from other workspaces (adopt_K)
            assert self.nq == 0
            check(lib().ab200_path_create_stage2(cat.handle, self.nf, self.np_, C.byref(self._h)))
        else:
            check(lib().ab200_path_create(cat.handle, self.nf, self.np_, self.nq, C.byref(self._h)))
        if stream is not None:
            check(lib().ab200_path_set_stream(self._h, C.c_void_p(int(stream))))

    def upload(self, f, atm: AtmPath, r, I_bkg, rte_option="linsrc", targets=(), select_species=abi.SPECIES_BATH,
               no_negative_absorption=1, hse_derivative=0, flags=0):
        f, stride, nf = _f_arg(f, atm.np_)
        assert nf == self.nf and atm.np_ <= self.np_cap, "the workspace takes paths of up to np_cap levels"
        self.np_ = atm.np_
        tg, _ = make_targets(targets)
        self._keep = (f, atm, np.ascontiguousarray(r, dtype=np.float64),
                      None if I_bkg is None else np.ascontiguousarray(I_bkg, dtype=np.float64))
        a = atm.desc()
        check(lib().ab200_path_upload(self._h, dptr(f), stride, C.byref(a), int(select_species),
                                      int(no_negative_absorption), tg, dptr(self._keep[2]), int(hse_derivative),
                                      _rte(rte_option), dptr(self._keep[3]), flags))

    def set_grid_bounds(self, bounds):
        """``bounds`` [np,2]: first / last frequency of the whole (unsharded) grid per level, or None."""
        b = None if bounds is None else np.ascontiguousarray(np.broadcast_to(np.asarray(bounds, float), (self.np_, 2)))
        check(lib().ab200_path_set_grid_bounds(self._h, dptr(b)))

    def set_timing(self, on=True):
        check(lib().ab200_path_set_timing(self._h, int(bool(on))))

    def timings(self):
        """{kernel class: (ms, launches)} since the last call; synchronises the path's stream."""
        ms = (C.c_double * 4)()
        n = (C.c_int64 * 4)()
        check(lib().ab200_path_get_timings(self._h, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(("prepare", "sum_real", "sum_cplx", "stokes"))}

    def region_histogram(self, samples_per_level=200_000, seed=1):
        """Sampled histogram of the reference's Faddeeva regions over this path's evaluations (see header)."""
        out = (C.c_double * 8)()
        check(lib().ab200_path_region_histogram(self._h, int(samples_per_level), int(seed), out))
        return np.array(out[:])

    def run_propmat(self):
        check(lib().ab200_path_run_propmat(self._h))

    def adopt_K(self):
        """The resident K was filled by the caller with rows summed by other workspaces of the same catalog / selection."""
        check(lib().ab200_path_adopt_K(self._h))

    def run_stokes(self):
        check(lib().ab200_path_run_stokes(self._h))

    def sync(self):
        check(lib().ab200_path_sync(self._h))

    def download(self, I=None, K=None, dI=None, dK=None):
        check(lib().ab200_path_download(self._h, dptr(I), dptr(dI), dptr(K), dptr(dK)))

    def add_lookup(self, lut: "Lookup", h2o_species=-1, target_d=(), p_interp_order=7, t_interp_order=7, water_interp_order=7,
                   f_interp_order=7, extpolfac=0.5, zero_init=True):
        """``spectral_propmatAddLookup`` (src/m_lookup.cc:143-173) on the resident K / dK; with ``zero_init`` instead of the
        line-by-line term (``spectral_propmat_agendaAuto(use_abs_lookup_data=1)``)."""
        d = np.ascontiguousarray(target_d, dtype=np.float64)
        check(lib().ab200_path_add_lookup(self._h, lut.handle, int(h2o_species), dptr(d if len(d) else None), int(p_interp_order),
                                          int(t_interp_order), int(water_interp_order), int(f_interp_order), float(extpolfac),
                                          int(bool(zero_init))))

    def add_predefined(self, models, species, target_d=(), data: "PredefData" = None):
        """``spectral_propmatAddPredefined`` (src/m_predefined_absorption_models.cc:156-191) on the resident K / dK: ``models`` are
        tag names ("O2-SelfContStandardType", ...), ``species`` maps "O2" / "N2" / "H2O" / "CO2" / "liquidcloud" to VMR indices,
        ``data`` carries the tables of the MT_CKD 4.x water continua."""
        ids, sp = abi.predef_args(models, species)
        d = np.ascontiguousarray(target_d, dtype=np.float64)
        check(lib().ab200_path_add_predefined_data(self._h, abi.ptr(ids, C.c_int32), len(ids), C.byref(sp), dptr(d if len(d) else None),
                                                   data.handle if data is not None else None))

    def add_cia(self, cia: "Cia", T_extrapolfac=0.5, ignore_errors=0, dT=0.1):
        """``spectral_propmatAddCIA`` (src/m_cia.cc:27-178) on the resident K / dK, after ``run_propmat``."""
        check(lib().ab200_path_add_cia(self._h, cia.handle, float(T_extrapolfac), int(ignore_errors), float(dT)))

    def run_observer(self, obs: abi.Observer):
        """Stokes chain + the host glue of ``spectral_rad_observer_agenda`` / ``measurement_vecFromSensor`` on the
        device: background from a temperature (src/m_background.cc:55-141), ``spectral_rad_jacFromBackground`` and
        ``spectral_rad_jacAddPathPropagation`` (src/m_rad.cc:26-127) inside the Jacobian pass,
        ``spectral_rad_transform_operator`` and ``SensorObsel::sumup`` (src/core/sensor/obsel.cpp:246-279)."""
        self._obs = obs
        d = obs.desc(self.np_, self.nq)
        check(lib().ab200_path_run_observer(self._h, C.byref(d)))

    def download_observer(self, want_jx=True):
        """Returns ``(spectral_rad [nf,4], spectral_rad_jac [nx,nf,4] or None, y [nch], Jy [nch,nx])``."""
        obs = self._obs
        I = np.empty((self.nf, 4))
        Jx = np.empty((obs.nx, self.nf, 4)) if want_jx and obs.nx else None
        nch = len(obs.channels)
        y = np.empty(nch) if nch else None
        Jy = np.empty((nch, obs.nx)) if nch and obs.nx else None
        check(lib().ab200_path_download_observer(self._h, dptr(I), dptr(Jx), dptr(y), dptr(Jy)))
        return I, Jx, y, Jy

    def device_ptr(self, which: int) -> int:
        return int(lib().ab200_path_device_ptr(self._h, which) or 0)

    def close(self):
        if self._h:
            lib().ab200_path_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measurement_vecFromSensor(abs_bands, freq_grid, simulations, jac_targets=(), rte_option="linsrc",
                              no_negative_absorption=1, hse_derivative=0, n_workspaces=2):
    """``measurement_vecFromSensor`` (src/m_rad.cc:301-362, the ``low_memory`` loop over ``SensorSimulations``) for
    clear-sky emission observers sharing one frequency grid: every simulation is one propagation path
    ``(atm_path, r, Observer[, I_bkg])`` (``I_bkg`` only when the observer has no background temperature); its spectral radiance and state-space Jacobian stay on the device and only its
    contribution to the channels comes back (``measurement_vec[iv] += obsel.sumup(...)``, :346-351).

    Paths may have different numbers of levels.  ``n_workspaces`` device workspaces on their own streams are used
    round-robin, so one path's transfers overlap the next one's kernels.  Returns ``(y [M], J [M, nx])``.
    """
    cat = _as_catalog(abs_bands)
    f = np.ascontiguousarray(freq_grid, dtype=np.float64)
    sims = list(simulations)
    if not sims:
        return np.zeros(0), np.zeros((0, 0))
    tg, nq = make_targets(jac_targets)
    M, nx = len(sims[0][2].channels), sims[0][2].nx
    np_cap = max(s[0].np_ for s in sims)
    y, J = np.zeros(M), np.zeros((M, nx))
    work = [Path(cat, len(f), np_cap, nq) for _ in range(max(1, min(n_workspaces, len(sims))))]
    busy = [False] * len(work)

    def drain(k):
        _, _, yk, Jk = work[k].download_observer(want_jx=False)
        if yk is not None:
            y[:] += yk
        if Jk is not None:
            J[:] += Jk
        busy[k] = False

    try:
        for i, sim in enumerate(sims):
            atm, r, obs = sim[:3]
            if obs.bkg_T is None and len(sim) < 4:
                raise ValueError("a simulation without background temperature needs its I_bkg")
            if len(obs.channels) != M or obs.nx != nx:
                raise ValueError("every simulation must address the same channels and state vector")
            k = i % len(work)
            if busy[k]:
                drain(k)
            w = work[k]
            w.upload(f, atm, r, None if obs.bkg_T is not None else sim[3], rte_option=rte_option,
                     targets=jac_targets, no_negative_absorption=no_negative_absorption, hse_derivative=hse_derivative)
            w.run_propmat()
            w.run_observer(obs)
            busy[k] = True
        n = len(sims)
        for i in range(max(0, n - len(work)), n):  # the rest in the order of the simulations: a fixed summation order
            if busy[i % len(work)]:
                drain(i % len(work))
    finally:
        for w in work:
            w.close()
    return y, J


def _host_catalog_from_desc(d) -> HostCatalog:
    """Copy of a loader's ``ab200_catalog_desc`` into numpy arrays (the loader's handle can be destroyed afterwards)."""
    nl, nls, nb, ni = d.n_lines, d.n_ls, d.n_bands, d.n_isot

    def arr(p, n, dtype):
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype)

    return HostCatalog(
        n_species=d.n_species, isot_species=arr(d.isot_species, ni, np.int32), isot_mass=arr(d.isot_mass, ni, np.float64),
        band_isot=arr(d.band_isot, nb, np.int32), band_offset=arr(d.band_offset, nb + 1, np.int64),
        f0=arr(d.f0, nl, np.float64), a=arr(d.a, nl, np.float64), e0=arr(d.e0, nl, np.float64),
        gu=arr(d.gu, nl, np.float64), gl=arr(d.gl, nl, np.float64), T0=arr(d.T0, nl, np.float64),
        ls_offset=arr(d.ls_offset, nl + 1, np.int64), ls_species=arr(d.ls_species, nls, np.int32),
        ls_type=arr(d.ls_type, nls * abi.NVAR, np.int32).reshape(nls, abi.NVAR),
        ls_X=arr(d.ls_X, nls * abi.NVAR * 4, np.float64).reshape(nls, abi.NVAR, 4),
        band_lineshape=arr(d.band_lineshape, nb, np.int32), band_cutoff_type=arr(d.band_cutoff_type, nb, np.int32),
        band_cutoff_value=arr(d.band_cutoff_value, nb, np.float64),
        z_on=arr(d.z_on, nl, np.uint8), z_gu=arr(d.z_gu, nl, np.float64), z_gl=arr(d.z_gl, nl, np.float64),
        two_Ju=arr(d.two_Ju, nl, np.int32), two_Jl=arr(d.two_Jl, nl, np.int32))


def abs_bandsReadXML(file=None, isotopologues=(), species_names=None, n_species=None, text=None) -> HostCatalog:
    """The reference's ``abs_bands`` XML (``xml_io_stream<AbsorptionBand>``, src/core/lbl/lbl_data.cpp:412-470) straight
    into the SoA catalog.  ``isotopologues``: ``(tag, species index, mass [g/mol])`` per SpeciesIsotope the file may
    name; ``species_names``: ``{name as written in the file: species index or SPECIES_BATH}`` for the broadeners.
    One band per ``<AbsorptionBand>`` in file order."""
    iso = (abi.XmlIsotopologue * len(isotopologues))()
    keep = []
    for k, (tag, sp, mass) in enumerate(isotopologues):
        keep.append(str(tag).encode())
        iso[k].name, iso[k].species, iso[k].mass = keep[-1], int(sp), float(mass)
    species_names = dict(species_names or {})
    nm = (abi.XmlSpecies * max(len(species_names), 1))()
    for k, (name, sp) in enumerate(species_names.items()):
        keep.append(str(name).encode())
        nm[k].name, nm[k].species = keep[-1], int(sp)
    ns = int(n_species) if n_species is not None else 1 + max([int(i[1]) for i in isotopologues] + [int(v) for v in species_names.values()])
    h = C.c_void_p()
    if text is not None:
        raw = text.encode() if isinstance(text, str) else bytes(text)
        check(lib().ab200_xml_read_bands(raw, len(raw), iso, len(isotopologues), nm, len(species_names), ns, C.byref(h)))
    else:
        check(lib().ab200_xml_read_bands_file(str(file).encode(), iso, len(isotopologues), nm, len(species_names), ns, C.byref(h)))
    try:
        return _host_catalog_from_desc(lib().ab200_xml_desc(h).contents)
    finally:
        lib().ab200_xml_destroy(h)


def abs_bandsReadHITRAN(file=None, frequency_range=(-np.inf, np.inf), isotopologues=(), n_species=None, text=None,
                        file_formatter=("par",), line_strength_option="S", compute_zeeman_parameters=0, n_threads=0):
    """``abs_bandsReadHITRAN`` (src/m_lbl.cc:302-338) for the plain 160-column ``.par`` format, straight into the SoA
    catalog: ``isotopologues`` is a list of ``(M, I, species index, mass [g/mol])`` (what ``Hitran::id_from_lookup`` and
    the isotopologue table give the shim), with two more entries ``(..., Hitran isotopologue ratio, Q(296 K))`` for
    ``line_strength_option="S"`` (the reference's default: the Einstein coefficient from the line strength, ``line::hitran_a``).
    Returns a ``HostCatalog`` (one band per isotopologue).

    Records below ``frequency_range[0]`` are skipped and reading stops at the first one above ``frequency_range[1]``
    (src/core/lbl/lbl_hitran.cpp:146-172).  Quantum-number columns and Zeeman parameters are outside this loader.
    """
    if tuple(file_formatter) != ("par",):
        raise Ab200Error(abi.ERR_UNSUPPORTED if hasattr(abi, "ERR_UNSUPPORTED") else 2,
                         "only file_formatter = ['par'] is on this path (no quantum-number columns)")
    if line_strength_option not in ("A", "S") or compute_zeeman_parameters:
        raise Ab200Error(2, "line_strength_option must be 'S' or 'A', and Zeeman parameters are outside this loader")
    opt = 0 if line_strength_option == "S" else 1  # AB200_HITRAN_STRENGTH_S / _A
    tab = (abi.HitranIsotopologue * len(isotopologues))()
    for k, row in enumerate(isotopologues):
        M, I, sp, mass = row[:4]
        tab[k].M, tab[k].I, tab[k].species, tab[k].mass = int(M), str(I).encode()[:1], int(sp), float(mass)
        if len(row) > 4:  # (..., hitran isotopologue ratio, Q(296 K)) for line_strength_option = "S"
            tab[k].hitran_ratio, tab[k].Q296 = float(row[4]), float(row[5])
    ns = int(n_species) if n_species is not None else 1 + max(int(i[2]) for i in isotopologues)
    h = C.c_void_p()
    if text is not None:
        raw = text.encode() if isinstance(text, str) else bytes(text)
        check(lib().ab200_hitran_read_par(raw, len(raw), float(frequency_range[0]), float(frequency_range[1]), opt, tab,
                                          len(isotopologues), ns, int(n_threads), C.byref(h)))
    else:
        check(lib().ab200_hitran_read_par_file(str(file).encode(), float(frequency_range[0]), float(frequency_range[1]), opt, tab,
                                               len(isotopologues), ns, int(n_threads), C.byref(h)))
    try:
        return _host_catalog_from_desc(lib().ab200_hitran_desc(h).contents)
    finally:
        lib().ab200_hitran_destroy(h)


class Cia:
    """``abs_cia_data`` on the device (ab200_cia): a list of ``_abi.CiaRecord``."""

    def __init__(self, records):
        self._records = list(records)
        arr = abi.cia_records(self._records)
        self._h = C.c_void_p()
        check(lib().ab200_cia_create(arr, len(self._records), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().ab200_cia_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spectral_propmatAddCIA(spectral_propmat, spectral_propmat_jac, freq_grid, jac_targets, select_species, abs_cia_data: Cia,
                           atm_path: AtmPath, T_extrapolfac=0.5, ignore_errors=0, dT=0.1):
    """src/m_cia.cc:27-178 for every level of ``atm_path``: ``spectral_propmat`` [np, nf, 7] and ``spectral_propmat_jac``
    [np, nq, nf, 7] are accumulated in place (``+=``).  ``dT`` is the perturbation of the temperature target."""
    np_ = atm_path.np_
    f, stride, nf = _f_arg(freq_grid, np_)
    tg, nq = make_targets(jac_targets)
    K = spectral_propmat
    if K.shape != (np_, nf, 7) or not K.flags.c_contiguous or K.dtype != np.float64:
        raise ValueError("*f_grid* must match *spectral_propmat*")
    dK = spectral_propmat_jac
    if nq and (dK is None or dK.shape != (np_, nq, nf, 7) or not dK.flags.c_contiguous):
        raise ValueError("*spectral_propmat_jac* must match derived form of *jac_targets*")
    a = atm_path.desc()
    check(lib().ab200_cia_levels(abs_cia_data.handle, nf, dptr(f), stride, C.byref(a), atm_path.vmr.shape[1], int(select_species), nq,
                                 tg, float(dT), float(T_extrapolfac), int(ignore_errors), dptr(K), dptr(dK if nq else None)))
    return K, dK


def partition_functions(tables, T):
    """``PartitionFunctions::Q`` / ``dQdT`` (src/partfun/partfun.h, generated by src/partfun/make_auto_partfuns.cc) for
    every level: ``tables[i] = (kind, grid | None, coef)`` with kind in interp / coeff / const / static_interp.
    Returns ``(Q [np, n_isot], dQdT [np, n_isot])`` ready for ``AtmPath``."""
    T = np.ascontiguousarray(T, dtype=np.float64)
    keep, arr = [], (abi.PartfunTable * len(tables))()
    for k, (kind, grid, coef) in enumerate(tables):
        g = None if grid is None else np.ascontiguousarray(grid, dtype=np.float64)
        c = np.ascontiguousarray(np.atleast_1d(coef), dtype=np.float64)
        keep.append((g, c))
        arr[k].kind, arr[k].n, arr[k].grid, arr[k].coef = abi.PARTFUN_KINDS[kind], len(c), dptr(g), dptr(c)
    Q = np.empty((len(T), len(tables)))
    dQ = np.empty((len(T), len(tables)))
    check(lib().ab200_partfun_eval(arr, len(tables), len(T), dptr(T), dptr(Q), dptr(dQ)))
    return Q, dQ


class AtmProfile:
    """The ``AtmField`` of a 1-D atmosphere flattened once (ab200_atm_profile): altitude grid, T, P, VMRs [nalt, ns], constant
    isotopologue ratios, optional magnetic field and wind [nalt, 3], the extrapolation rule below / above the grid
    (``"None" | "Zero" | "Nearest" | "Linear"``), ``top_of_atmosphere`` and the partition-function tables of the
    isotopologues (as for ``partition_functions``)."""

    def __init__(self, alt, T, P, vmr, isorat, mag=None, wind=None, alt_low="None", alt_upp="None", top_of_atmosphere=None,
                 partfun_tables=None):
        a = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.float64)  # noqa: E731
        self.alt, self.T, self.P, self.vmr, self.isorat, self.mag, self.wind = a(alt), a(T), a(P), a(vmr), a(isorat), a(mag), a(wind)
        self.vmr = self.vmr.reshape(len(self.alt), -1)
        self.alt_low, self.alt_upp = alt_low, alt_upp
        self.toa = float(self.alt[-1] if top_of_atmosphere is None else top_of_atmosphere)
        self._pf_keep, self._pf = [], None
        if partfun_tables is not None:
            self._pf = (abi.PartfunTable * len(partfun_tables))()
            for k, (kind, grid, coef) in enumerate(partfun_tables):
                g = None if grid is None else np.ascontiguousarray(grid, dtype=np.float64)
                c = np.ascontiguousarray(np.atleast_1d(coef), dtype=np.float64)
                self._pf_keep.append((g, c))
                self._pf[k].kind, self._pf[k].n, self._pf[k].grid, self._pf[k].coef = abi.PARTFUN_KINDS[kind], len(c), dptr(g), dptr(c)

    def desc(self):
        d = abi.AtmProfileDesc()
        d.nalt, d.alt, d.T, d.P, d.vmr, d.isorat = len(self.alt), dptr(self.alt), dptr(self.T), dptr(self.P), dptr(self.vmr), dptr(self.isorat)
        d.mag, d.wind = dptr(self.mag), dptr(self.wind)
        d.alt_low, d.alt_upp, d.top_of_atmosphere = abi.EXTRAP[self.alt_low], abi.EXTRAP[self.alt_upp], self.toa
        d.partfun = self._pf
        return d


def atm_pathFromPath(ray_path_alt, atm_field: AtmProfile, los=None, in_atm=None) -> AtmPath:
    """``atm_pathFromPath`` (src/m_ppvar.cc:38-45) for a 1-D atmosphere: ``AtmField::at`` at every path point (the top of the
    atmosphere for points outside it), as flat arrays ready for the path entry points.  ``ray_path_alt`` = pos[0] of the
    points; ``los`` [np, 2] is carried along."""
    alt = np.ascontiguousarray(ray_path_alt, dtype=np.float64)
    n, ns, ni = len(alt), atm_field.vmr.shape[1], len(atm_field.isorat)
    T, P, vmr, iso = np.empty(n), np.empty(n), np.empty((n, ns)), np.empty((n, ni))
    has_pf = atm_field._pf is not None  # without tables Q stays NaN: the caller fills it (PartitionFunctions::Q)
    Q, dQ = (np.empty((n, ni)), np.empty((n, ni))) if has_pf else (np.full((n, ni), np.nan), None)
    mag = np.empty((n, 3)) if atm_field.mag is not None else None
    wind = np.empty((n, 3)) if atm_field.wind is not None else None
    ia = None if in_atm is None else np.ascontiguousarray(in_atm, dtype=np.uint8)
    d = atm_field.desc()
    check(lib().ab200_atm_path_from_profile(C.byref(d), ns, ni, n, dptr(alt), None if ia is None else ia.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            dptr(T), dptr(P), dptr(vmr), dptr(iso), dptr(Q if has_pf else None), dptr(dQ), dptr(mag), dptr(wind)))
    return AtmPath(T=T, P=P, vmr=vmr, isorat=iso, Q=Q, dQdT=dQ, mag=mag, los=None if los is None else np.asarray(los, dtype=np.float64),
                   wind=wind)


class Lookup:
    """``abs_lookup_data`` on the device (ab200_lookup): a list of ``_abi.LookupTable``."""

    def __init__(self, tables):
        self._tables = list(tables)
        arr = abi.lookup_tables(self._tables)
        self._h = C.c_void_p()
        check(lib().ab200_lookup_create(arr, len(self._tables), C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().ab200_lookup_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def abs_lookup_dataPrecompute(abs_bands, atm_profile: AtmPath, freq_grid, select_species, temperature_perturbation=None,
                              water_perturbation=None, h2o_species=None, partfun_tables=None) -> abi.LookupTable:
    """``abs_lookup_dataPrecompute`` (src/m_lookup.cc:175-197, the ``lookup::table`` constructor src/core/lookup/lookup_map.cpp:22-131)
    with the line-by-line sum on the GPU: for every temperature offset and water ratio the reference profile is perturbed and
    ``lbl::calculate`` runs for the selected species with ``no_negative_absorption = true``; all nt * nw * np levels go through
    the device in one launch sequence per (offset, ratio) (``ab200_lookup_precompute``).  ``xsec = K.A / number_density(species)``."""
    cat = _as_catalog(abs_bands)
    f = np.ascontiguousarray(freq_grid, dtype=np.float64)
    tp = [0.0] if temperature_perturbation is None else list(temperature_perturbation)
    wp = [1.0] if water_perturbation is None else list(water_perturbation)
    if water_perturbation is not None and h2o_species is None:
        raise ValueError("a water perturbation grid needs the index of H2O in the VMR vector")
    if atm_profile.np_ > 1 and not (np.diff(atm_profile.P) < 0).all():
        raise ValueError("the reference profile of a lookup table must have descending pressures (DescendingGrid log_p_grid, "
                         "src/core/lookup/lookup_map.h)")
    xsec = np.empty((len(tp), len(wp), atm_profile.np_, len(f)))
    tpa = None if temperature_perturbation is None else np.ascontiguousarray(tp, dtype=np.float64)
    wpa = None if water_perturbation is None else np.ascontiguousarray(wp, dtype=np.float64)
    a = atm_profile.desc()
    pf, keep = None, []
    if partfun_tables is not None:  # Q at the perturbed temperatures (the reference evaluates it inside lbl::calculate)
        pf = (abi.PartfunTable * len(partfun_tables))()
        for k, (kind, grid, coef) in enumerate(partfun_tables):
            g = None if grid is None else np.ascontiguousarray(grid, dtype=np.float64)
            cf = np.ascontiguousarray(np.atleast_1d(coef), dtype=np.float64)
            keep.append((g, cf))
            pf[k].kind, pf[k].n, pf[k].grid, pf[k].coef = abi.PARTFUN_KINDS[kind], len(cf), dptr(g), dptr(cf)
    check(lib().ab200_lookup_precompute(cat.handle, len(f), dptr(f), C.byref(a), int(select_species),
                                        -1 if h2o_species is None else int(h2o_species), len(tp), dptr(tpa), len(wp), dptr(wpa),
                                        pf, dptr(xsec)))
    return abi.LookupTable(species=int(select_species), f_grid=f, log_p_grid=np.log(atm_profile.P), t_atmref=np.array(atm_profile.T, float),
                           xsec=xsec, t_pert=None if temperature_perturbation is None else np.asarray(tp, float),
                           w_pert=None if water_perturbation is None else np.asarray(wp, float),
                           water_atmref=None if water_perturbation is None else np.array(atm_profile.vmr[:, h2o_species], float))


def spectral_propmatAddLookup(spectral_propmat, spectral_propmat_jac, freq_grid, jac_targets, select_species, abs_lookup_data: Lookup,
                              atm_path: AtmPath, h2o_species=-1, target_d=(), no_negative_absorption=1, p_interp_order=7,
                              t_interp_order=7, water_interp_order=7, f_interp_order=7, extpolfac=0.5):
    """src/m_lookup.cc:143-173 for every level of ``atm_path``: ``spectral_propmat`` [np, nf, 7] is accumulated, the rows of
    ``spectral_propmat_jac`` [np, nq, nf, 7] are assigned (like the reference, :130-135)."""
    np_ = atm_path.np_
    f, stride, nf = _f_arg(freq_grid, np_)
    tg, nq = make_targets(jac_targets)
    K, dK = spectral_propmat, spectral_propmat_jac
    if K.shape != (np_, nf, 7) or not K.flags.c_contiguous or K.dtype != np.float64:
        raise ValueError("spectral_propmat must be a C-contiguous float64 [np, nf, 7] array")
    d = np.ascontiguousarray(target_d, dtype=np.float64)
    a = atm_path.desc()
    check(lib().ab200_lookup_levels(abs_lookup_data.handle, nf, dptr(f), stride, C.byref(a), atm_path.vmr.shape[1], int(h2o_species),
                                    int(select_species), nq, tg, dptr(d if nq else None), int(no_negative_absorption), int(p_interp_order),
                                    int(t_interp_order), int(water_interp_order), int(f_interp_order), float(extpolfac), dptr(K),
                                    dptr(dK if nq else None)))
    return K, dK


class PredefData:
    """The data part of ``abs_predef_data`` on the device (ab200_predef_data): what ``abs_predef_dataAddWaterMTCKD400`` / ``...430``
    (src/m_predefined_absorption_models.cc:69-148) load.  ``ckdmt400`` / ``ckdmt430``: dicts with ref_temp [K], ref_press [hPa],
    wavenumbers, self_absco_ref, for_absco_ref, self_texp (equal lengths >= 4, ascending regular wavenumbers), or None."""

    def __init__(self, ckdmt400=None, ckdmt430=None, device=0):
        keep = []

        def desc(w):
            if w is None:
                return None
            cols = [np.ascontiguousarray(w[k], dtype=np.float64) for k in ("wavenumbers", "self_absco_ref", "for_absco_ref", "self_texp")]
            if len({len(c) for c in cols}) != 1:
                raise ValueError("Mismatching size, all vector inputs must be the same length")
            keep.extend(cols)
            return abi.MtckdWater(len(cols[0]), float(w["ref_temp"]), float(w["ref_press"]), *(dptr(c) for c in cols))

        a, b = desc(ckdmt400), desc(ckdmt430)
        self._h = C.c_void_p()
        check(lib().ab200_predef_data_create(C.byref(a) if a is not None else None, C.byref(b) if b is not None else None, int(device),
                                             C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().ab200_predef_data_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spectral_propmatAddPredefined(spectral_propmat, spectral_propmat_jac, abs_predef_data, select_species, jac_targets, freq_grid,
                                  atm_path: AtmPath, species, target_d=(), data: PredefData = None):
    """src/m_predefined_absorption_models.cc:156-191 for every level: ``abs_predef_data`` is a list of model tag names (``data``
    their tables, MT_CKD 4.x only), ``species`` maps "O2" / "N2" / "H2O" / "CO2" / "liquidcloud" to VMR indices; K [np, nf, 7]
    and dK [np, nq, nf, 7] are +=."""
    np_ = atm_path.np_
    f, stride, nf = _f_arg(freq_grid, np_)
    tg, nq = make_targets(jac_targets)
    ids, sp = abi.predef_args(abs_predef_data, species)
    K, dK = spectral_propmat, spectral_propmat_jac
    if K.shape != (np_, nf, 7) or not K.flags.c_contiguous or K.dtype != np.float64:
        raise ValueError("Mismatch dimensions on internal matrices of xsec and frequency")
    d = np.ascontiguousarray(target_d, dtype=np.float64)
    a = atm_path.desc()
    check(lib().ab200_predef_levels_data(abi.ptr(ids, C.c_int32), len(ids), C.byref(sp), nf, dptr(f), stride, C.byref(a),
                                         atm_path.vmr.shape[1], int(select_species), nq, tg, dptr(d if nq else None), dptr(K),
                                         dptr(dK if nq else None), data.handle if data is not None else None))
    return K, dK
