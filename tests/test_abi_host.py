"""CPU-side checks of the drop-in boundary and the host logic (no GPU, no compute calls).

* the C-ABI shared library loads and exports every symbol include/arts_b200.h declares;
* without a CUDA device every compute entry point fails loudly (no CPU fallback);
* the host-only pieces of the library — Zeeman sub-line pre-expansion (closed-form 3j for j2 = 1,
  reference lbl_zeeman.cpp:261-309 via wigner3j) and zeeman::norm_view (:413-455) — against the
  oracle's independent restatement (Racah-formula 3j);
* the synthetic configs are deterministic and have the sizes BASELINE.json names.
"""
import ctypes as C
import re

import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import _lib, shard, synth
from arts_b200._abi import dptr


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 25, names
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/arts_b200.h but not exported by libarts_b200.so"
    for must in ("ab200_catalog_create", "ab200_propmat_levels", "ab200_tramat", "ab200_srcvec", "ab200_rte_emission",
                 "ab200_clearsky_emission", "ab200_last_error", "ab200_path_set_grid_bounds"):
        assert must in names


def test_header_has_no_cxx_or_torch_types():
    src = open(_lib.HEADER_PATH).read()
    body = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for bad in ("std::", "torch", "at::", "template", "class ", "&"):
        assert bad not in body, bad
    assert 'extern "C"' in body


def test_python_struct_mirrors_match_header_field_order():
    src = open(_lib.HEADER_PATH).read()
    body = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    m = re.search(r"typedef struct ab200_catalog_desc \{(.*?)\} ab200_catalog_desc;", body, re.S)
    fields = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        fields += [re.sub(r"[\s\*]", "", n) for n in re.sub(r"^(const\s+)?\w+\s", "", decl, count=1).split(",")]
    assert fields == [n for n, _ in abi.CatalogDesc._fields_]
    m = re.search(r"typedef struct ab200_atm_path \{(.*?)\} ab200_atm_path;", body, re.S)
    fields = [re.sub(r"[\s\*]", "", re.sub(r"^(const\s+)?\w+\s", "", d.strip(), count=1)) for d in m.group(1).split(";")
              if d.strip()]
    assert fields == [n for n, _ in abi.AtmPathDesc._fields_]


@pytest.mark.skipif(_lib.lib().ab200_device_count() > 0, reason="only meaningful without a CUDA device")
def test_no_device_means_loud_failure_not_fallback():
    from arts_b200 import wsm

    c = synth.tiny_case(nl=8, nf=16, np_=2)
    with pytest.raises(_lib.Ab200Error) as e:
        wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    assert e.value.code == abi.ERR_CUDA
    with pytest.raises(_lib.Ab200Error):
        wsm.faddeeva_w(np.array([1 + 1j]))
    with pytest.raises(_lib.Ab200Error):
        wsm.spectral_tramat_pathFromPath(np.zeros((2, 4, 7)), None, np.ones(1), np.full(2, 250.0))


@pytest.mark.parametrize("pol", [1, 2, 3])
def test_zeeman_components_closed_form_vs_wigner(orc, pol):
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for tJl in range(0, 12):
        for tJu in (tJl - 2, tJl, tJl + 2):
            if tJu < 0:
                continue
            gu, gl = rng.uniform(-2, 2, 2)
            s_ref, d_ref = orc.zeeman_components(1, gu, gl, tJu, tJl, pol)
            cap = 64
            s, d = np.zeros(cap), np.zeros(cap)
            n = L.ab200_zeeman_components(1, gu, gl, tJu, tJl, pol, cap, dptr(s), dptr(d))
            assert n == len(s_ref) == tJl + 1
            np.testing.assert_allclose(s[:n], s_ref, rtol=1e-13, atol=1e-16)
            np.testing.assert_allclose(d[:n], d_ref, rtol=1e-14, atol=0)
            if tJu == tJl and tJl == 0:
                assert not s[:n].any()  # J = 0 -> 0 is forbidden


def test_zeeman_strengths_sum_rule(orc):
    """Sum over the sub-lines of (pi + sigma- + sigma+) = 3 for an allowed transition... per
    polarisation the relative strengths of lbl_zeeman.cpp:261-277 sum to 1 (x 0.75 / 1.5 factors)."""
    L = _lib.lib()
    for tJl, tJu in ((2, 4), (4, 2), (6, 6), (3, 5), (1, 1)):
        tot = []
        for pol in (1, 2, 3):
            s, d = np.zeros(64), np.zeros(64)
            n = L.ab200_zeeman_components(1, 1.0, 1.0, tJu, tJl, pol, 64, dptr(s), dptr(d))
            tot.append(s[:n].sum())
        np.testing.assert_allclose(tot, [0.5, 0.25, 0.25], rtol=1e-13)


def test_norm_view_vs_oracle(orc):
    L = _lib.lib()
    rng = np.random.default_rng(1)
    for _ in range(200):
        mag = rng.normal(0, 40e-6, 3)
        los = np.array([rng.uniform(0, 180), rng.uniform(-180, 180)])
        for pol in range(4):
            out = np.empty(7)
            L.ab200_norm_view(pol, dptr(mag), dptr(los), dptr(out))
            np.testing.assert_allclose(out, orc.norm_view(pol, mag, los), rtol=1e-13, atol=1e-15)
    out = np.empty(7)
    L.ab200_norm_view(0, dptr(np.zeros(3)), dptr(np.zeros(2)), dptr(out))
    assert list(out) == [1, 0, 0, 0, 0, 0, 0]


def test_synthetic_configs_are_deterministic_and_sized():
    a, b = synth.case_c1(), synth.case_c1()
    assert a.n_lines == 1000 and a.nf == 10_000 and a.np_ == 1
    assert np.array_equal(a.cat.f0, b.cat.f0) and np.array_equal(a.cat.ls_X, b.cat.ls_X)
    assert np.all(np.diff(a.cat.f0) >= 0)
    c2 = synth.case_c2(lines_per_species=200, nf=1000, np_=100)
    assert c2.cat.n_species == 5 and c2.np_ == 100 and len(c2.r) == 99
    assert c2.cat.n_lines == 1000 and len(c2.cat.band_isot) == 100
    c3 = synth.case_c3(nf=38 * 10, np_=50)
    assert c3.cat.n_lines == 38 and c3.cat.z_on.all() and c3.np_ == 50
    assert np.all(np.diff(c3.f) >= 0)
    c4 = synth.case_c4(n_lines=1000, nf=10_000, np_=100, f_slice=(2500, 5000))
    full = synth.case_c4(n_lines=1000, nf=10_000, np_=100)
    assert np.array_equal(c4.f, full.f[2500:5000]) and np.array_equal(c4.cat.f0, full.cat.f0)


def test_frequency_ranges_mirror_omp_offset_count():
    """matpack::omp_offset_count (matpack_mdspan_algorithm.cc:4-18): nf // n per block, remainder to the last."""
    for nf, n in ((10, 1), (10, 3), (1_000_000, 8), (7, 7), (1001, 4), (5, 2)):
        rs = shard.frequency_ranges(nf, n)
        assert len(rs) == n and rs[0][0] == 0
        assert sum(c for _, c in rs) == nf
        for (o0, c0), (o1, _) in zip(rs, rs[1:]):
            assert o1 == o0 + c0 and c0 == nf // n
    with pytest.raises(ValueError):
        shard.frequency_ranges(10, 0)


def test_host_catalog_desc_keeps_arrays_alive_and_typed():
    c = synth.tiny_case(nl=16, nf=8, np_=2)
    d = c.cat.desc()
    assert d.n_lines == c.cat.n_lines and d.n_ls == len(c.cat.ls_species)
    assert d.f0[0] == c.cat.f0[0] and d.band_offset[d.n_bands] == d.n_lines
    a = c.atm.desc()
    assert a.np == 2 and a.T[1] == c.atm.T[1]
    assert C.sizeof(abi.Target) == 24  # kind, species, line (int64), ls_var, coeff


def test_header_compiles_as_c99_and_host_entry_points_work_from_c(tmp_path):
    """The boundary is a C ABI: include/arts_b200.h must be plain C (a cgo / JNI / N-API stub includes it as such), and the
    host-only entry points (HITRAN ingest, partition functions, error reporting) are exercised from a C program linked
    against libarts_b200.so (tests/c/abi_smoke.c)."""
    import os
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "arts_b200")
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
           os.path.join(root, "tests", "c", "abi_smoke.c"), "-o", exe, "-L" + lib_dir, "-larts_b200", "-lm", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "abi smoke ok" in r.stdout, r.stdout + r.stderr
