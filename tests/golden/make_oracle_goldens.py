"""Writes the oracle-made golden vectors of tests/golden/ (run in the CPU container, commit the output).

    python tests/golden/make_oracle_goldens.py

* linsrc_convergence.json   brightness temperatures of the reference's catalog-free fixture
                            tests/core/linsrc/test_linsrc_convergence.py (inputs fully specified there)
* path_goldens.npz          propagation matrices / radiances of small seeded synthetic cases
                            (arts_b200.synth) from oracle/_ref/liboracle.so, which links the reference's
                            own Faddeeva.cc object.  The GPU parity tests compare with these on the box
                            in addition to the live oracle.
* jacobian_goldens.npz      propagation-matrix and radiance Jacobians of a small Zeeman path, one target of every
                            kind (temperature, VMR, wind, magnetic field, isotopologue ratio, line centre, line-shape
                            coefficient)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from arts_b200 import synth  # noqa: E402
from tests import oracle_lib as orc  # noqa: E402
from tests.test_oracle_pins import _linsrc_fixture  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_cases():
    """name -> Case; shared with tests/test_gpu_goldens.py."""
    return {
        "c1_small": synth.case_c1(nl=200, nf=1200),
        "c1_cutoff": synth.case_c1(nl=200, nf=1200, cutoff=1.5e9),
        "c2_small": synth.case_c2(lines_per_species=100, nf=900, np_=12, bands_per_species=2),
        "c3_small": synth.case_c3(nf=38 * 12, np_=9, los=(120.0, 30.0)),
    }


def jacobian_case():
    """Small Zeeman path and one target of every kind the library takes; shared with tests/test_gpu_goldens.py."""
    from arts_b200 import _abi as abi

    c = synth.case_c3(nf=38 * 4, np_=4, los=(130.0, 25.0))
    line = 5
    tg = (("T",), ("VMR", 0), ("wind_w",), ("mag_u",), ("isorat", 0), ("line_f0", line),
          ("line_ls", line, abi.VAR_G0, abi.SPECIES_BATH, 0))
    return c, tg


def main():
    out = {"constant_k": _linsrc_fixture(orc, False), "varying": _linsrc_fixture(orc, True),
           "source": "tests/core/linsrc/test_linsrc_convergence.py:24-92,110-178 through oracle/oracle.cpp"}
    json.dump(out, open(os.path.join(HERE, "linsrc_convergence.json"), "w"), indent=1)
    arrays = {}
    for name, c in golden_cases().items():
        K, _ = orc.propmat_levels(c.cat, c.f, c.atm)
        arrays[name + "_K"] = K
        if c.np_ > 1:
            for opt in ("constant", "linsrc"):
                I, _ = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option=opt)
                arrays[f"{name}_I_{opt}"] = I
                arrays[f"{name}_Tb_{opt}"] = orc.planck_tb(c.f, I)
    np.savez_compressed(os.path.join(HERE, "path_goldens.npz"), **arrays)
    c, tg = jacobian_case()
    K, dK = orc.propmat_levels(c.cat, c.f, c.atm, targets=tg)
    I, dI = orc.clearsky_emission(c.cat, c.f, c.atm, c.r, c.I_bkg, targets=tg, hse_derivative=1)
    np.savez_compressed(os.path.join(HERE, "jacobian_goldens.npz"), K=K, dK=dK, I=I, dI=dI)
    print("wrote", sorted(arrays), "threads", orc.num_threads())


if __name__ == "__main__":
    main()
