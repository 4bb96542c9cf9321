"""Catalog ingest: the HITRAN .par loader of the C ABI (host code, no GPU needed) against the
pure-Python restatement of the reference's reader (tests/hitran_ref.py) and the reference's own fixture
tests/hitran/single_line.par (tests/golden/hitran_single_line.json)."""
import json
import os

import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import wsm
from tests import hitran_ref

HERE = os.path.dirname(os.path.abspath(__file__))
# (M, I, species, mass, HITRAN isotopologue ratio, Q(296 K)); the last two only matter for line_strength_option = "S"
TABLE = [(1, "1", 0, 18.010565, 0.997317, 174.58), (1, "2", 0, 20.014811, 1.99983e-3, 176.05), (2, "1", 1, 43.98983, 0.984204, 286.09),
         (7, "1", 2, 31.98983, 0.995262, 215.73), (7, "A", 2, 35.0, 1e-3, 500.0)]


def _read(*a, **k):
    """The loader with line_strength_option = "A" (the file's own Einstein coefficient) unless a test says otherwise."""
    k.setdefault("line_strength_option", "A")
    return wsm.abs_bandsReadHITRAN(*a, **k)


def _record(M, I, nu, S, A, ga, gs, E, n, d, gu, gl, filler=None):
    def fortran(x, width, digits):  # Fw.d drops the leading zero when the field is too narrow for it
        t = f"{x:.{digits}f}"
        if len(t) > width:
            t = t.replace("0.", ".", 1)
        return t.rjust(width)

    body = f"{M:2d}{I}{nu:12.6f}{S:10.3E}{A:10.3E}{fortran(ga, 5, 4)}{gs:5.3f}{E:10.4f}{n:4.2f}{fortran(d, 8, 6)}"
    assert len(body) == 67, (len(body), body)
    filler = (filler or "").ljust(79)[:79]
    return body + filler + f"{gu:7.1f}{gl:7.1f}"


def _synthetic_file(n=3000, seed=4):
    rng = np.random.default_rng(seed)
    nu = np.sort(rng.uniform(0.05, 4000.0, n))
    recs = []
    for k in range(n):
        M, I = [(1, "1"), (1, "2"), (2, "1"), (7, "1"), (7, "A")][int(rng.integers(5))]
        d = 0.0 if k % 7 == 0 else float(rng.uniform(-0.02, 0.02))
        ga = float(rng.uniform(0.01, 0.12))
        recs.append(_record(M, I, nu[k], 10 ** rng.uniform(-30, -19), 10 ** rng.uniform(-12, 2), ga, float(rng.uniform(0.05, 0.9)),
                            float(rng.uniform(0, 5000)), float(rng.uniform(0.3, 0.9)), d, float(rng.integers(1, 60)),
                            float(rng.integers(1, 60)), filler="          0 1 0          0 0 0  4  2  2        5  1  5"))
    return "\n".join(recs) + "\n"


def _compare(cat, ref):
    assert len(cat.f0) == len(ref)
    order = []
    for i in range(len(TABLE)):  # one band per isotopologue with lines, table order, file order inside
        order += [k for k, r in enumerate(ref) if r["isot"] == i]
    bands = [i for i in range(len(TABLE)) if any(r["isot"] == i for r in ref)]
    assert list(cat.band_isot) == bands
    assert list(np.diff(cat.band_offset)) == [sum(r["isot"] == i for r in ref) for i in bands]
    for name, key in (("f0", "f0"), ("a", "a"), ("e0", "e0"), ("gu", "gu"), ("gl", "gl")):
        assert np.array_equal(getattr(cat, name), np.array([ref[k][key] for k in order])), name  # bit-exact
    assert (cat.T0 == 296.0).all()
    assert np.array_equal(cat.ls_offset, 2 * np.arange(len(ref) + 1))
    for j, k in enumerate(order):
        r = ref[k]
        assert cat.ls_species[2 * j] == TABLE[r["isot"]][2] and cat.ls_species[2 * j + 1] == abi.SPECIES_BATH
        for e, g in ((2 * j, r["gamma_self"]), (2 * j + 1, r["gamma_air"])):
            assert cat.ls_type[e, abi.VAR_G0] == abi.TM_T1
            assert cat.ls_X[e, abi.VAR_G0, 0] == g and cat.ls_X[e, abi.VAR_G0, 1] == r["n"]
            if r["delta"] != 0:
                assert cat.ls_type[e, abi.VAR_D0] == abi.TM_T0 and cat.ls_X[e, abi.VAR_D0, 0] == r["delta"]
            else:
                assert cat.ls_type[e, abi.VAR_D0] == abi.TM_ABSENT
            assert (cat.ls_type[e, [abi.VAR_DV, abi.VAR_Y, abi.VAR_G]] == abi.TM_ABSENT).all()


def test_reference_fixture_single_line():
    """The reference's fixture: the 2.16 GHz H2O line, read column by column."""
    g = json.load(open(os.path.join(HERE, "golden", "hitran_single_line.json")))
    col = g["columns"]
    cat = _read(text=g["record"] + "\n", isotopologues=TABLE)
    c100 = 100 * 299792458.0
    assert len(cat.f0) == 1 and cat.band_isot[0] == 0 and cat.isot_species[0] == 0
    assert cat.f0[0] == col["nu_cm-1"] * c100
    assert abs(cat.f0[0] - 2.15997e9) < 1e5
    assert cat.a[0] == col["A_s-1"] and cat.gu[0] == col["g_upp"] and cat.gl[0] == col["g_low"]
    assert cat.e0[0] == col["E_cm-1"] * (6.62607015e-34 * c100)
    k = (1 / 101325.0) * c100
    assert cat.ls_X[0, abi.VAR_G0, 0] == col["gamma_self_cm-1_atm-1"] * k and cat.ls_X[1, abi.VAR_G0, 0] == col["gamma_air_cm-1_atm-1"] * k
    assert cat.ls_X[0, abi.VAR_G0, 1] == col["n_air"] and cat.ls_X[0, abi.VAR_D0, 0] == col["delta_cm-1_atm-1"] * k
    ref = hitran_ref.read_par(g["record"] + "\n", -np.inf, np.inf, TABLE)
    _compare(cat, ref)
    # with the quantum-number tail the record is longer than the `par` format: the reference's error
    with pytest.raises(wsm.Ab200Error, match="Part of the line was not parsed"):
        _read(text=g["record"] + ",ElecStateLabel=X\n", isotopologues=TABLE)


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_synthetic_file_matches_restatement(threads, tmp_path):
    text = _synthetic_file()
    for (fmin, fmax) in ((-np.inf, np.inf), (3e12, 6e13), (5e13, 5.0001e13), (2e14, 3e14)):
        ref = hitran_ref.read_par(text, fmin, fmax, TABLE)
        cat = _read(text=text, frequency_range=(fmin, fmax), isotopologues=TABLE, n_threads=threads)
        _compare(cat, ref)
    path = tmp_path / "lines.par"
    path.write_text(text)
    cat_f = _read(file=str(path), isotopologues=TABLE, n_threads=threads)
    _compare(cat_f, hitran_ref.read_par(text, -np.inf, np.inf, TABLE))
    # no trailing newline
    _compare(_read(text=text[:-1], isotopologues=TABLE), hitran_ref.read_par(text[:-1], -np.inf, np.inf, TABLE))


def test_crlf_and_one_trailing_separator_load_like_the_reference():
    """`if (not data.end_of_string()) data.skip(1)` (lbl_hitran.cpp:126): a 161-byte record — a CRLF file, or one trailing
    separator — loads; only from 162 bytes on is the remainder an error (:133-135)."""
    recs = [_record(1, "1", 100.0 + 7 * i, 1e-25, 1e-3, 0.07, 0.4, 300.0, 0.7, -0.004, 9, 11) for i in range(5)]
    unix = "\n".join(recs) + "\n"
    ref = hitran_ref.read_par(unix, -np.inf, np.inf, TABLE)
    for text in ("\r\n".join(recs) + "\r\n", ",\n".join(recs) + ",\n", "\r\n".join(recs)):
        _compare(_read(text=text, isotopologues=TABLE), ref)
        _compare(_read(text=text, isotopologues=TABLE), hitran_ref.read_par(text, -np.inf, np.inf, TABLE))
    with pytest.raises(wsm.Ab200Error, match="Part of the line was not parsed: 'x'"):
        _read(text=recs[0] + ",x\n", isotopologues=TABLE)


def test_error_behaviour_follows_the_reference():
    good = _record(1, "1", 100.0, 1e-25, 1e-3, 0.07, 0.4, 300.0, 0.7, -0.004, 9, 11)
    above = _record(1, "1", 3000.0, 1e-25, 1e-3, 0.07, 0.4, 300.0, 0.7, -0.004, 9, 11)
    cases = {
        "unknown isotopologue": (_record(9, "1", 100.0, 1e-25, 1e-3, 0.07, 0.4, 300.0, 0.7, 0.0, 9, 11), "isotopologue table"),
        "short record": (good[:120], "Unexpected end of string"),
        "two trailing characters": (good + " \r", "Part of the line was not parsed"),
        "garbage in a column": (good[:25] + "  1.0E-0x " + good[35:], "Failed to parse value"),
        "zero upper degeneracy": (_record(1, "1", 100.0, 1e-25, 1e-3, 0.07, 0.4, 300.0, 0.7, 0.0, 0, 11), "Invalid Einstein coefficient"),
        "blank line": ("", "Unexpected end of string"),
    }
    for name, (bad, msg) in cases.items():
        text = good + "\n" + bad + "\n" + good + "\n"
        with pytest.raises(wsm.Ab200Error, match=msg):
            _read(text=text, isotopologues=TABLE)
        with pytest.raises(hitran_ref.HitranError):
            hitran_ref.read_par(text, -np.inf, np.inf, TABLE)
        # the same record below the window is never looked at beyond its frequency (:72-73) ...
        if name not in ("short record", "blank line", "two trailing characters"):
            cat = _read(text=text, frequency_range=(kay(100.5), np.inf), isotopologues=TABLE)
            assert len(cat.f0) == 0
        # ... and one after the first record above the window is never read (:163-166)
        text2 = good + "\n" + above + "\n" + bad + "\n"
        cat = _read(text=text2, frequency_range=(-np.inf, kay(2000.0)), isotopologues=TABLE)
        assert len(cat.f0) == 1
        assert len(hitran_ref.read_par(text2, -np.inf, kay(2000.0), TABLE)) == 1
    with pytest.raises(wsm.Ab200Error, match="Cannot open file"):
        _read(file="/nonexistent/lines.par", isotopologues=TABLE)
    with pytest.raises(wsm.Ab200Error):
        _read(text=good + "\n", isotopologues=TABLE, file_formatter=("par", "statep", "statepp"))


def kay(x):
    return x * 100 * 299792458.0


@pytest.mark.gpu
def test_hitran_catalog_through_the_gpu_path(orc):
    """The loaded SoA is what ab200_catalog_create consumes: propagation matrix against the oracle on the same arrays."""
    text = _synthetic_file(n=800, seed=9)
    cat = _read(text=text, frequency_range=(kay(500.0), kay(900.0)), isotopologues=TABLE)
    assert 50 < len(cat.f0) < 400
    cat.a *= 1e-3
    nlev = 3
    atm = abi.AtmPath(T=np.array([290.0, 250.0, 220.0]), P=np.array([9e4, 3e4, 5e3]),
                      vmr=np.tile([1e-2, 4e-4, 0.21], (nlev, 1)), isorat=np.tile([0.997, 0.002, 0.984, 0.995, 0.001], (nlev, 1)),
                      Q=np.tile([170.0, 180.0, 280.0, 215.0, 230.0], (nlev, 1)) * (np.array([290.0, 250.0, 220.0]) / 296.0)[:, None],
                      dQdT=np.tile([0.6, 0.6, 0.9, 0.7, 0.8], (nlev, 1)))
    f = np.linspace(kay(480.0), kay(920.0), 900)
    tg = (("T",), ("VMR", 0))
    Kr, dKr = orc.propmat_levels(cat, f, atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(cat, f, atm, jac_targets=tg)
    from tests.conftest import assert_propmat_close
    from tests.test_gpu_jacobian import assert_jac_close
    assert_propmat_close(K, Kr)
    for q in range(2):
        assert_jac_close(dK[:, q], dKr[:, q], what=f"HITRAN catalog dK target {q}")
    assert Kr[..., 0].max() > 0


def test_partition_function_tables():
    """PartitionFunctions::Q / dQdT for the four table kinds of src/partfun/make_auto_partfuns.cc:28-153, against a pure-Python
    restatement of the generated code (bit-exact) and against numpy (interp / polyval) for the meaning."""
    import bisect

    rng = np.random.default_rng(6)
    grid = np.sort(rng.uniform(50.0, 600.0, 12))
    qv = 10.0 * (grid / 296.0) ** 1.5
    sgrid = np.arange(100.0, 401.0, 25.0)
    sq = 7.0 * (sgrid / 296.0) ** 1.5
    poly = np.array([1.5, 0.3, 2e-3, -1e-6])
    tables = [("interp", grid, qv), ("coeff", None, poly), ("const", None, [42.0]), ("static_interp", sgrid, sq)]
    T = np.concatenate([rng.uniform(20.0, 700.0, 40), grid[[0, 3, -1]], sgrid[[0, 5, -1]], [99.9, 400.1]])
    Q, dQ = wsm.partition_functions(tables, T)

    def ref(kind, g, c, t):
        if kind == "interp":  # :46-61
            i_low = bisect.bisect_left(list(g), t)
            i = min(i_low - (i_low > 0), len(c) - 2)
            return c[i] + (t - g[i]) * (c[i + 1] - c[i]) / (g[i + 1] - g[i]), (c[i + 1] - c[i]) / (g[i + 1] - g[i])
        if kind == "coeff":  # :77-103
            res, TN = c[0], 1.0
            for k in range(1, len(c)):
                TN *= t
                res += TN * c[k]
            d, TN = c[1], 1.0
            for k in range(2, len(c)):
                TN *= t
                d += float(k) * TN * c[k]
            return res, d
        if kind == "const":
            return c[0], 0.0
        r_dT = 1.0 / (g[1] - g[0])  # :117-153
        Tx = (t - g[0]) * r_dT
        iTx = int(Tx) if Tx >= 0 else 2 ** 63  # static_cast<Size> of a negative double: huge, clamped below
        i = len(c) - 2 if iTx > len(c) - 2 else iTx
        To = Tx - float(i)
        return c[i] + To * (c[i + 1] - c[i]), (c[i + 1] - c[i]) * r_dT

    for k, (kind, g, c) in enumerate(tables):
        for j, t in enumerate(T):
            if kind == "static_interp" and t < g[0]:
                continue  # the conversion of a negative double to size_t is undefined behaviour in the generated code
            q, d = ref(kind, g, np.atleast_1d(np.asarray(c, float)), float(t))
            assert Q[j, k] == q and dQ[j, k] == d, (kind, t)
    inside = (T >= grid[0]) & (T <= grid[-1])
    np.testing.assert_allclose(Q[inside, 0], np.interp(T[inside], grid, qv), rtol=1e-14)
    np.testing.assert_allclose(Q[:, 1], np.polyval(poly[::-1], T), rtol=1e-13)
    np.testing.assert_allclose(dQ[:, 1], np.polyval(np.polyder(poly[::-1]), T), rtol=1e-13)
    ins = (T >= sgrid[0]) & (T <= sgrid[-1])
    np.testing.assert_allclose(Q[ins, 3], np.interp(T[ins], sgrid, sq), rtol=1e-13)
    assert (Q[:, 2] == 42.0).all() and not dQ[:, 2].any()
    with pytest.raises(wsm.Ab200Error, match="Temperature grid must be increasing"):
        wsm.partition_functions([("interp", grid[::-1], qv)], T)


def test_line_strength_option_S_is_the_default_and_matches_the_file(tmp_path):
    """HitranLineStrengthOption::S (the reference's default, src/workspace_methods.cpp:3278): the Einstein coefficient
    comes from the line-strength column, a = einstein_a(S / ratio, gu, e0, f0, 296 K, Q(296)) (line::hitran_a,
    lbl_data.cpp:34-40,155-169).  Bit for bit against the Python restatement; on the reference's fixture record the
    result reproduces the file's own A column (HITRAN computes one from the other) within 2e-4; g_upp = 0 -> gu = gl = -1."""
    g = json.load(open(os.path.join(HERE, "golden", "hitran_single_line.json")))
    text = g["record"] + "\n"
    cat = wsm.abs_bandsReadHITRAN(text=text, isotopologues=TABLE)  # default option
    ref = hitran_ref.read_par(text, -np.inf, np.inf, TABLE, option="S")
    assert cat.a[0] == ref[0]["a"]
    assert cat.a[0] == pytest.approx(g["columns"]["A_s-1"], rel=2e-4)
    cat_a = _read(text=text, isotopologues=TABLE)
    assert cat_a.a[0] == g["columns"]["A_s-1"] and cat_a.f0[0] == cat.f0[0] and cat_a.e0[0] == cat.e0[0]
    rng = np.random.default_rng(9)
    lines = []
    for i in range(200):
        M, I = [(1, "1"), (1, "2"), (2, "1"), (7, "1")][i % 4]
        lines.append(_record(M, I, 10.0 + 5.0 * i, float(10 ** rng.uniform(-28, -20)), float(10 ** rng.uniform(-8, 0)), 0.07, 0.3,
                             float(rng.uniform(0, 3000)), 0.7, -0.002, 0.0 if i % 50 == 7 else float(2 * (i % 9) + 1), 3.0))
    text = "\n".join(lines) + "\n"
    cat = wsm.abs_bandsReadHITRAN(text=text, isotopologues=TABLE, n_threads=3)
    ref = hitran_ref.read_par(text, -np.inf, np.inf, TABLE, option="S")
    _compare(cat, ref)
    assert (cat.gu == -1.0).sum() == 4 and (cat.gl[cat.gu == -1.0] == -1.0).all()
    with pytest.raises(wsm.Ab200Error):  # option S without the two extra table entries
        wsm.abs_bandsReadHITRAN(text=text, isotopologues=[t[:4] for t in TABLE])
