"""ctypes binding of the CPU oracle (oracle/_ref/liboracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module.  The oracle takes the same flattened C-ABI structs as the
CUDA library (arts_b200._abi), so both sides see byte-identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from arts_b200 import _abi as abi
from arts_b200._abi import AtmPath, HostCatalog, dptr, make_targets

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_ref", "liboracle.so")
_lib = None


def build(force=False):
    """Build oracle/_ref/liboracle.so when the reference tree is present (CPU container)."""
    ref = os.environ.get("ARTS_REFERENCE", "/root/reference")
    if os.path.isdir(ref):
        cmd = ["make", "-C", os.path.join(_ROOT, "oracle"), f"REF={ref}"] + (["-B"] if force else [])
        subprocess.run(cmd, check=True, capture_output=True)
    if not os.path.exists(_SO):
        raise RuntimeError(f"{_SO} is missing and {ref} is not available to build it")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.orc_last_error.restype = C.c_char_p
        L.orc_faddeeva_w.argtypes = [C.c_int64, dp, dp, dp, dp]
        L.orc_propmat_levels.argtypes = [C.POINTER(abi.CatalogDesc)] + abi.SIG_PROPMAT_LEVELS_CORE + [dp, dp]
        L.orc_tramat.argtypes = abi.SIG_TRAMAT
        L.orc_srcvec.argtypes = abi.SIG_SRCVEC
        L.orc_rte_emission.argtypes = abi.SIG_RTE
        L.orc_rte_transmission.argtypes = abi.SIG_TRANSMISSION
        L.orc_clearsky_emission.argtypes = [C.POINTER(abi.CatalogDesc)] + abi.SIG_CLEARSKY_CORE
        L.orc_planck_tb.argtypes = [C.c_int64, dp, dp]
        L.orc_planck.argtypes = [C.c_int64, dp, C.c_double, dp]
        L.orc_tran.argtypes = [dp, dp, C.c_double, C.c_uint32, dp, dp]
        L.orc_sqrt_propmat.argtypes = [dp, dp]
        L.orc_dawson.argtypes = [C.c_int64, dp, dp, dp, dp]
        L.orc_wigner3j.argtypes = [C.c_int] * 6 + [dp]
        L.orc_wind_shift.argtypes = [dp, dp, dp, dp]
        L.orc_zeeman_components.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int64, dp, dp]
        L.orc_norm_view.argtypes = [C.c_int, dp, dp, dp]
        L.orc_dnorm_view.argtypes = [C.c_int, C.c_int, dp, dp, dp]
        L.orc_cia_levels.argtypes = [C.POINTER(abi.CiaRecordDesc), C.c_int32, C.c_int64, dp, C.c_int64, C.POINTER(abi.AtmPathDesc),
                                     C.c_int32, C.c_int32, C.c_int32, C.POINTER(abi.Target), C.c_double, C.c_double, C.c_int32, dp, dp]
        L.orc_lookup_levels.argtypes = [C.POINTER(abi.LookupTableDesc), C.c_int32, C.c_int64, dp, C.c_int64, C.POINTER(abi.AtmPathDesc),
                                        C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(abi.Target), dp, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_double, dp, dp]
        L.orc_predef_levels.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(abi.PredefSpecies), C.c_int64, dp, C.c_int64,
                                        C.POINTER(abi.AtmPathDesc), C.c_int32, C.c_int32, C.c_int32, C.POINTER(abi.Target), dp, dp, dp]
        L.orc_predef_levels_data.argtypes = L.orc_predef_levels.argtypes + [C.POINTER(abi.MtckdWater), C.POINTER(abi.MtckdWater)]
        L.orc_background.argtypes = [C.c_int64, dp, C.c_double, dp, dp]
        L.orc_observer.argtypes = [C.c_int32, C.c_int64, C.c_int32, dp, C.POINTER(abi.ObserverDesc), dp, dp, dp, dp, dp, dp]
        for name in ("orc_invplanck", "orc_dinvplanckdI", "orc_invrayjean", "orc_dplanck_dt"):
            getattr(L, name).argtypes = [C.c_double, C.c_double]
            getattr(L, name).restype = C.c_double
        _lib = L
    return _lib


def _check(rc):
    if rc:
        raise RuntimeError(f"oracle error {rc}: {lib().orc_last_error().decode()}")


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP team size of the oracle (bench.py's CPU arm: all host cores, whatever OMP_NUM_THREADS the launcher exported)."""
    return int(lib().orc_set_num_threads(int(n)))


def faddeeva_w(z):
    z = np.ascontiguousarray(z, dtype=np.complex128).ravel()
    zr, zi = np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)
    wr, wi = np.empty_like(zr), np.empty_like(zr)
    _check(lib().orc_faddeeva_w(len(zr), dptr(zr), dptr(zi), dptr(wr), dptr(wi)))
    return wr + 1j * wi


def dawson(z):
    """the reference's own Faddeeva::Dawson(complex) object code"""
    z = np.ascontiguousarray(z, dtype=np.complex128).ravel()
    zr, zi = np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)
    dr, di = np.empty_like(zr), np.empty_like(zr)
    _check(lib().orc_dawson(len(zr), dptr(zr), dptr(zi), dptr(dr), dptr(di)))
    return dr + 1j * di


def _f_arg(f, np_):
    f = np.ascontiguousarray(f, dtype=np.float64)
    if f.ndim == 1:
        return f, 0, f.shape[0]
    assert f.shape[0] == np_
    return f, f.shape[1], f.shape[1]


def propmat_levels(cat: HostCatalog, f, atm: AtmPath, select_species=abi.SPECIES_BATH, no_negative_absorption=1,
                   targets=(), K=None, dK=None):
    f, stride, nf = _f_arg(f, atm.np_)
    tg, nq = make_targets(targets)
    K = np.zeros((atm.np_, nf, 7)) if K is None else K
    dK = np.zeros((atm.np_, nq, nf, 7)) if dK is None else dK
    d, a = cat.desc(), atm.desc()
    _check(lib().orc_propmat_levels(C.byref(d), nf, dptr(f), stride, C.byref(a), select_species,
                                    no_negative_absorption, nq, tg, dptr(K), dptr(dK)))
    return K, dK


def wind_shift(wind, los):
    """(fac, freq_wind_shift_jac[3]) of one path point."""
    wind = np.ascontiguousarray(wind, dtype=np.float64)
    los = np.ascontiguousarray(los, dtype=np.float64)
    fac, jac = np.zeros(1), np.zeros(3)
    _check(lib().orc_wind_shift(dptr(wind), dptr(los), dptr(fac), dptr(jac)))
    return float(fac[0]), jac


def tramat(K, dK, r, dr, rte_option, flags=0):
    np_, nf, _ = K.shape
    nq = dK.shape[1] if dK is not None and dK.size else 0
    K = np.ascontiguousarray(K)
    dK = np.zeros((np_, 0, nf, 7)) if nq == 0 else np.ascontiguousarray(dK)
    r = np.ascontiguousarray(r, dtype=np.float64)
    dr = np.zeros((2, np_ - 1, nq)) if dr is None else np.ascontiguousarray(dr, dtype=np.float64)
    T = np.empty((nf, np_, 16))
    L = np.empty((nf, np_, 16))
    P = np.empty((nf, np_, 16))
    dT = np.empty((2, nf, np_, nq, 16))
    dL = np.empty((2, nf, np_, nq, 16))
    _check(lib().orc_tramat(np_, nf, nq, dptr(K), dptr(dK), dptr(r), dptr(dr), abi.RTE_OPTIONS[rte_option], flags,
                            dptr(T), dptr(L), dptr(P), dptr(dT), dptr(dL)))
    return T, L, P, dT, dL


def srcvec(K, f, T_level, it=-1, nq=0):
    np_, nf, _ = K.shape
    f, stride, _ = _f_arg(f, np_)
    T_level = np.ascontiguousarray(T_level, dtype=np.float64)
    J = np.empty((nf, np_, 4))
    dJ = np.empty((nf, np_, nq, 4))
    _check(lib().orc_srcvec(np_, nf, nq, dptr(np.ascontiguousarray(K)), dptr(f), stride, dptr(T_level), it, dptr(J),
                            dptr(dJ)))
    return J, dJ


def rte_emission(rte_option, T, L, P, dT, dL, J, dJ, I_bkg):
    nf, np_, _ = T.shape
    nq = dJ.shape[2]
    I_bkg = np.ascontiguousarray(I_bkg, dtype=np.float64)
    I = np.empty((nf, 4))
    dI = np.empty((nf, np_, nq, 4))
    _check(lib().orc_rte_emission(abi.RTE_OPTIONS[rte_option], np_, nf, nq, dptr(T), dptr(L), dptr(P), dptr(dT),
                                  dptr(dL), dptr(J), dptr(dJ), dptr(I_bkg), dptr(I), dptr(dI)))
    return I, dI


def rte_transmission(T, P, dT, I_bkg):
    """rte_transmission (rtepack_rtestep.cc:456-503): (I [nf,4], dI [nf,np,nq,4])."""
    nf, np_, _ = P.shape
    nq = 0 if dT is None else dT.shape[3]
    I_bkg = np.ascontiguousarray(I_bkg, dtype=np.float64)
    I = np.empty((nf, 4))
    dI = np.zeros((nf, np_, nq, 4))
    _check(lib().orc_rte_transmission(np_, nf, nq, dptr(np.ascontiguousarray(T)), dptr(np.ascontiguousarray(P)),
                                      dptr(None if dT is None else np.ascontiguousarray(dT)), dptr(I_bkg), dptr(I), dptr(dI)))
    return I, dI


def clearsky_emission(cat, f, atm, r, I_bkg, rte_option="linsrc", select_species=abi.SPECIES_BATH,
                      no_negative_absorption=1, targets=(), hse_derivative=0, flags=0, return_K=False):
    f, stride, nf = _f_arg(f, atm.np_)
    tg, nq = make_targets(targets)
    r = np.ascontiguousarray(r, dtype=np.float64)
    I_bkg = np.ascontiguousarray(I_bkg, dtype=np.float64)
    I = np.empty((nf, 4))
    dI = np.empty((nf, atm.np_, nq, 4))
    K = np.empty((atm.np_, nf, 7)) if return_K else None
    d, a = cat.desc(), atm.desc()
    _check(lib().orc_clearsky_emission(C.byref(d), nf, dptr(f), stride, C.byref(a), select_species,
                                       no_negative_absorption, nq, tg, dptr(r), hse_derivative,
                                       abi.RTE_OPTIONS[rte_option], dptr(I_bkg), flags, dptr(I), dptr(dI), dptr(K)))
    return (I, dI, K) if return_K else (I, dI)


def planck_tb(f, I):
    f = np.ascontiguousarray(f, dtype=np.float64)
    out = np.array(I, dtype=np.float64, order="C", copy=True)
    _check(lib().orc_planck_tb(len(f), dptr(f), dptr(out)))
    return out


def planck(f, T):
    f = np.ascontiguousarray(f, dtype=np.float64)
    B = np.empty_like(f)
    _check(lib().orc_planck(len(f), dptr(f), float(T), dptr(B)))
    return B


def tran(k1, k2, r, flags=0, linsrc=True):
    k1 = np.ascontiguousarray(k1, dtype=np.float64)
    k2 = np.ascontiguousarray(k2, dtype=np.float64)
    T = np.empty(16)
    L = np.empty(16)
    _check(lib().orc_tran(dptr(k1), dptr(k2), float(r), flags, dptr(T), dptr(L) if linsrc else dptr(None)))
    return T.reshape(4, 4), L.reshape(4, 4)


def wigner3j(tj1, tj2, tj3, tm1, tm2, tm3):
    out = C.c_double()
    lib().orc_wigner3j(tj1, tj2, tj3, tm1, tm2, tm3, C.byref(out))
    return out.value


def zeeman_components(on, gu, gl, tJu, tJl, pol):
    cap = 4 * (tJl + 2)
    s = np.zeros(cap)
    d = np.zeros(cap)
    n = lib().orc_zeeman_components(int(on), gu, gl, tJu, tJl, pol, cap, dptr(s), dptr(d))
    return s[:n], d[:n]


def norm_view(pol, mag, los):
    mag = np.ascontiguousarray(mag, dtype=np.float64)
    los = np.ascontiguousarray(los, dtype=np.float64)
    out = np.empty(7)
    lib().orc_norm_view(pol, dptr(mag), dptr(los), dptr(out))
    return out


def dnorm_view(pol, comp, mag, los):
    mag = np.ascontiguousarray(mag, dtype=np.float64)
    los = np.ascontiguousarray(los, dtype=np.float64)
    out = np.empty(7)
    lib().orc_dnorm_view(pol, comp, dptr(mag), dptr(los), dptr(out))
    return out


def background(f, T):
    """from_temp (src/m_background.cc:55-63): (I_bkg [nf,4], dB/dT [nf])."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    I = np.empty((len(f), 4))
    dB = np.empty(len(f))
    _check(lib().orc_background(len(f), dptr(f), float(T), dptr(I), dptr(dB)))
    return I, dB


def observer(f, obs: abi.Observer, P, I, dI):
    """Steps 2-5 of the observer epilogue: returns (I transformed [nf,4], Jx [nx,nf,4], y [nch], Jy [nch,nx])."""
    f = np.ascontiguousarray(f, dtype=np.float64)
    nf = len(f)
    P = np.ascontiguousarray(P, dtype=np.float64)
    np_ = P.shape[1]
    dI = None if dI is None else np.ascontiguousarray(dI, dtype=np.float64)
    nq = 0 if dI is None else dI.shape[2]
    d = obs.desc(np_, nq)
    Io = np.array(I, dtype=np.float64, order="C", copy=True)
    Jx = np.zeros((obs.nx, nf, 4))
    y = np.zeros(d.n_channels)
    Jy = np.zeros((d.n_channels, obs.nx))
    _check(lib().orc_observer(np_, nf, nq, dptr(f), C.byref(d), dptr(P), dptr(Io), dptr(dI), dptr(Jx), dptr(y), dptr(Jy)))
    return Io, Jx, y, Jy


def invplanck(i, f):
    return lib().orc_invplanck(float(i), float(f))


def dinvplanckdI(i, f):
    return lib().orc_dinvplanckdI(float(i), float(f))


def invrayjean(i, f):
    return lib().orc_invrayjean(float(i), float(f))


def dplanck_dt(f, t):
    return lib().orc_dplanck_dt(float(f), float(t))


def cia_levels(records, f, atm: AtmPath, select_species=abi.SPECIES_BATH, targets=(), dT=0.1, T_extrapolfac=0.5, ignore_errors=0,
               K=None, dK=None):
    """spectral_propmatAddCIA (src/m_cia.cc:27-178) per level; returns (K [np,nf,7], dK [np,nq,nf,7]) accumulated."""
    f, stride, nf = _f_arg(f, atm.np_)
    tg, nq = make_targets(targets)
    K = np.zeros((atm.np_, nf, 7)) if K is None else K
    dK = np.zeros((atm.np_, nq, nf, 7)) if dK is None else dK
    arr = abi.cia_records(records)
    a = atm.desc()
    _check(lib().orc_cia_levels(arr, len(records), nf, dptr(f), stride, C.byref(a), atm.vmr.shape[1], select_species, nq, tg,
                                float(dT), float(T_extrapolfac), int(ignore_errors), dptr(K), dptr(dK)))
    return K, dK


def lookup_levels(tables, f, atm: AtmPath, h2o_species=-1, select_species=abi.SPECIES_BATH, targets=(), target_d=(),
                  no_negative_absorption=1, orders=(7, 7, 7, 7), extpolfac=0.5, K=None, dK=None):
    """_spectral_propmatAddLookup (src/m_lookup.cc:20-141) per level; orders = (p, t, water, f)."""
    f, stride, nf = _f_arg(f, atm.np_)
    tg, nq = make_targets(targets)
    K = np.zeros((atm.np_, nf, 7)) if K is None else K
    dK = np.zeros((atm.np_, nq, nf, 7)) if dK is None else dK
    arr = abi.lookup_tables(tables)
    d = np.ascontiguousarray(target_d, dtype=np.float64)
    a = atm.desc()
    _check(lib().orc_lookup_levels(arr, len(tables), nf, dptr(f), stride, C.byref(a), atm.vmr.shape[1], int(h2o_species), select_species,
                                   nq, tg, dptr(d), int(no_negative_absorption), *[int(o) for o in orders], float(extpolfac), dptr(K),
                                   dptr(dK)))
    return K, dK


def _mtckd(w, keep):
    if w is None:
        return None
    cols = [np.ascontiguousarray(w[k], dtype=np.float64) for k in ("wavenumbers", "self_absco_ref", "for_absco_ref", "self_texp")]
    keep.extend(cols)
    return C.byref(abi.MtckdWater(len(cols[0]), float(w["ref_temp"]), float(w["ref_press"]), *(dptr(c) for c in cols)))


def predef_levels(models, species, f, atm: AtmPath, select_species=abi.SPECIES_BATH, targets=(), target_d=(), K=None, dK=None,
                  ckdmt400=None, ckdmt430=None):
    """spectral_propmatAddPredefined per level (m_predefined_absorption_models.cc:156-191); ckdmt400 / ckdmt430: the MT_CKD 4.x
    water tables as dicts (arts_b200.wsm.PredefData)."""
    f, stride, nf = _f_arg(f, atm.np_)
    tg, nq = make_targets(targets)
    ids, sp = abi.predef_args(models, species)
    K = np.zeros((atm.np_, nf, 7)) if K is None else K
    dK = np.zeros((atm.np_, nq, nf, 7)) if dK is None else dK
    d = np.ascontiguousarray(target_d, dtype=np.float64)
    a = atm.desc()
    keep = []
    _check(lib().orc_predef_levels_data(abi.ptr(ids, C.c_int32), len(ids), C.byref(sp), nf, dptr(f), stride, C.byref(a), atm.vmr.shape[1],
                                        select_species, nq, tg, dptr(d), dptr(K), dptr(dK), _mtckd(ckdmt400, keep), _mtckd(ckdmt430, keep)))
    return K, dK
