"""The GPU entry points of the C ABI called from a plain C program (tests/c/abi_gpu.c), the way the reference-side shim
of INTEGRATION.md calls them: no Python, no ctypes between the caller and libarts_b200.so.  The program links the CPU
oracle as its checker."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_entry_points_from_c(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    from tests import oracle_lib as orc

    orc.lib()  # oracle/_ref/liboracle.so exists (built here when the reference is present, shipped to the GPU box otherwise)
    lib_dir, orc_dir = os.path.join(ROOT, "arts_b200"), os.path.join(ROOT, "oracle", "_ref")
    exe = str(tmp_path / "abi_gpu")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_gpu.c"), "-o", exe, "-L" + lib_dir, "-larts_b200", "-L" + orc_dir, "-loracle", "-lm",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + orc_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "abi gpu ok" in r.stdout, f"exit {r.returncode}: {r.stdout}{r.stderr}"
