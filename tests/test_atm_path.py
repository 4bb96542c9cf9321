"""atm_pathFromPath for a 1-D atmosphere (src/m_ppvar.cc:38-45 -> forward_atm_path, atm_path.cpp:19-28 ->
Atm::Field::at, atm_field.cpp:890-947).  Host function: runs without a GPU.  Checked bitwise against the oracle's
restatement of the reference's lag / limit code and against an independent numpy formulation (np.interp) to 1e-13; every
InterpolationExtrapolation rule, the top-of-atmosphere substitution for points outside the atmosphere, and the reference's
error cases."""
import ctypes as C

import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import wsm
from arts_b200._abi import dptr
from tests import oracle_lib as orc


def _profile(rng, nalt=37, ns=3, ni=2, **kw):
    alt = np.sort(rng.uniform(0.0, 90e3, nalt))
    alt[0] = 0.0
    T = 288.0 - 60.0 * np.sin(alt / 30e3) + rng.normal(size=nalt)
    P = 101325.0 * np.exp(-alt / 7e3)
    vmr = np.abs(rng.normal(size=(nalt, ns))) * 1e-3
    iso = np.array([0.995, 0.004])[:ni]
    mag = rng.normal(size=(nalt, 3)) * 3e-5
    wind = rng.normal(size=(nalt, 3)) * 20.0
    pf = [("interp", np.linspace(100.0, 400.0, 31), 200.0 * (np.linspace(100.0, 400.0, 31) / 296.0) ** 1.5), ("coeff", None, [3.0, 0.7, 1e-4])][:ni]
    return wsm.AtmProfile(alt, T, P, vmr, iso, mag=mag, wind=wind, partfun_tables=pf, **kw)


def _oracle(prof, alt, in_atm=None):
    L = orc.lib()
    dp = C.POINTER(C.c_double)
    L.orc_atm_path_from_profile.argtypes = [C.POINTER(abi.AtmProfileDesc), C.c_int32, C.c_int32, C.c_int32, dp, C.POINTER(C.c_uint8)] + [dp] * 6
    n, ns, ni = len(alt), prof.vmr.shape[1], len(prof.isorat)
    T, P, vmr, iso, mag, wind = np.empty(n), np.empty(n), np.empty((n, ns)), np.empty((n, ni)), np.empty((n, 3)), np.empty((n, 3))
    ia = None if in_atm is None else np.ascontiguousarray(in_atm, dtype=np.uint8).ctypes.data_as(C.POINTER(C.c_uint8))
    d = prof.desc()
    orc._check(L.orc_atm_path_from_profile(C.byref(d), ns, ni, n, dptr(np.ascontiguousarray(alt)), ia, dptr(T), dptr(P), dptr(vmr), dptr(iso),
                                           dptr(mag), dptr(wind)))
    return T, P, vmr, iso, mag, wind


def test_inside_the_grid_bitwise_vs_oracle_and_numpy():
    rng = np.random.default_rng(21)
    prof = _profile(rng)
    alt = np.concatenate([rng.uniform(0, prof.alt[-1], 400), prof.alt, prof.alt[:-1] + 1e-9])
    a = wsm.atm_pathFromPath(alt, prof)
    T, P, vmr, iso, mag, wind = _oracle(prof, alt)
    for got, want in ((a.T, T), (a.P, P), (a.vmr, vmr), (a.isorat, iso), (a.mag, mag), (a.wind, wind)):
        assert np.array_equal(got, want)
    np.testing.assert_allclose(a.T, np.interp(alt, prof.alt, prof.T), rtol=1e-13)
    np.testing.assert_allclose(a.P, np.interp(alt, prof.alt, prof.P), rtol=1e-12)
    for s in range(prof.vmr.shape[1]):
        np.testing.assert_allclose(a.vmr[:, s], np.interp(alt, prof.alt, prof.vmr[:, s]), rtol=1e-12, atol=1e-18)
    Q, dQ = wsm.partition_functions([("interp", np.linspace(100.0, 400.0, 31), 200.0 * (np.linspace(100.0, 400.0, 31) / 296.0) ** 1.5),
                                     ("coeff", None, [3.0, 0.7, 1e-4])], a.T)
    assert np.array_equal(a.Q, Q) and np.array_equal(a.dQdT, dQ)  # PartitionFunctions::Q at the path temperatures


@pytest.mark.parametrize("rule", ["Zero", "Nearest", "Linear"])
def test_extrapolation_rules(rule):
    rng = np.random.default_rng(22)
    prof = _profile(rng, alt_low=rule, alt_upp=rule, top_of_atmosphere=120e3)
    prof.alt[0] = 500.0  # a grid that starts above the ground
    prof.vmr[:] = 1e-3 * (1.0 + np.outer(prof.alt, [1e-6, 2e-6, -3e-6]))  # stays positive under linear extrapolation
    alt = np.array([0.0, 499.0, 500.0, 60e3, prof.alt[-1], prof.alt[-1] + 1.0, 119e3])
    a = wsm.atm_pathFromPath(alt, prof)
    T, P, vmr, iso, mag, wind = _oracle(prof, alt)
    assert np.array_equal(a.T, T) and np.array_equal(a.P, P) and np.array_equal(a.vmr, vmr) and np.array_equal(a.mag, mag)
    out = (alt < prof.alt[0]) | (alt > prof.alt[-1])
    if rule == "Zero":
        assert np.all(a.T[out] == 0.0) and np.all(a.vmr[out] == 0.0)
    elif rule == "Nearest":
        assert a.T[0] == prof.T[0] and a.T[-1] == prof.T[-1]
    else:
        slope = (prof.T[1] - prof.T[0]) / (prof.alt[1] - prof.alt[0])
        assert a.T[0] == pytest.approx(prof.T[0] + slope * (0.0 - prof.alt[0]), rel=1e-12)
    assert np.all(a.T[~out] == np.interp(alt[~out], prof.alt, prof.T)) or np.allclose(a.T[~out], np.interp(alt[~out], prof.alt, prof.T), rtol=1e-13)


def test_points_outside_the_atmosphere_take_the_top():
    rng = np.random.default_rng(23)
    prof = _profile(rng, alt_upp="Nearest", top_of_atmosphere=100e3)
    alt = np.array([400e3, 30e3, 800e3])  # a satellite, a point inside, a point behind
    a = wsm.atm_pathFromPath(alt, prof, in_atm=[0, 1, 0])
    top = wsm.atm_pathFromPath([100e3], prof)
    assert a.T[0] == top.T[0] == a.T[2] and a.P[0] == top.P[0]
    assert np.array_equal(a.T, _oracle(prof, alt, in_atm=[0, 1, 0])[0])


def test_single_level_field_and_error_cases():
    rng = np.random.default_rng(24)
    one = wsm.AtmProfile([0.0], [250.0], [1e4], [[0.21, 1e-3]], [1.0], top_of_atmosphere=50e3)
    a = wsm.atm_pathFromPath([0.0, 10e3, 49e3], one)
    assert np.all(a.T == 250.0) and np.all(a.vmr[:, 0] == 0.21) and np.isnan(a.Q).all() and a.mag is None
    prof = _profile(rng)  # extrapolation None
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.atm_pathFromPath([prof.alt[-1] + 10.0], wsm.AtmProfile(prof.alt, prof.T, prof.P, prof.vmr, prof.isorat, top_of_atmosphere=200e3))
    assert "Limit breached" in str(e.value)
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.atm_pathFromPath([prof.alt[-1] + 10.0], prof)
    assert "above the top of the atmosphere" in str(e.value)
    bad = _profile(rng)
    bad.vmr[5, 1] = -1e-3
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.atm_pathFromPath(bad.alt[4:7] + 1.0, bad)
    assert "VMR" in str(e.value)
    bad.T[3] = np.nan
    with pytest.raises(wsm.Ab200Error) as e:
        wsm.atm_pathFromPath(bad.alt[2:4], bad)
    assert "Temperature is NaN" in str(e.value)
