#!/usr/bin/env python
"""Debug helper: dK of one fuzz case with and without the per-pair closed form of the Jacobian kernel, against the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from arts_b200 import wsm  # noqa: E402
from tests import oracle_lib as orc  # noqa: E402
from tests.test_gpu_fuzz import random_case  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 9
cat, f, atm, r, bkg = random_case(seed)
if len(sys.argv) > 2 and sys.argv[2] == "nocut":
    cat.band_cutoff_type[:] = 0
    cat.band_cutoff_value[:] = np.inf
tg = (("T",), ("VMR", 0))
Kr, dKr = orc.propmat_levels(cat, f, atm, no_negative_absorption=seed % 2, targets=tg)
out = {}
for flag in ("1", "0"):
    os.environ["AB200_JAC_PAIR_FAR"] = flag
    K, dK = wsm.spectral_propmat_pathFromPath(cat, f, atm, jac_targets=tg, no_negative_absorption=seed % 2)
    out[flag] = dK
    for q in range(2):
        sc = np.abs(dKr[:, q, :, 0]).max()
        err = np.abs(dK[:, q, :, 0] - dKr[:, q, :, 0])
        i = np.unravel_index(np.argmax(err), err.shape)
        print(f"pair_far={flag} target {q}: worst {err.max() / sc:.3e} at level {i[0]} freq {i[1]} f={f[i[1]]:.6e} "
              f"gpu {dK[i[0], q, i[1], 0]:.6e} ref {dKr[i[0], q, i[1], 0]:.6e}")
d = np.abs(out["1"] - out["0"])[..., 0]
i = np.unravel_index(np.argmax(d), d.shape)
print("largest difference between the two at", i, d.max(), "n_bands", len(cat.band_isot), "cutoffs", cat.band_cutoff_value,
      "band_isot", cat.band_isot, "isot_species", cat.isot_species)
lev = i[0]
print("T", atm.T[lev], "P", atm.P[lev], "vmr", atm.vmr[lev])
k = np.argsort(np.abs(cat.f0 - f[i[2]]))[:5]
print("nearest lines", cat.f0[k], "f", f[i[2]])
for q in range(2):
    print("q", q, "at worst point: pair", out["1"][i[0], q, i[2], 0], "generic", out["0"][i[0], q, i[2], 0], "oracle", dKr[i[0], q, i[2], 0],
          "max |dK| of this target", np.abs(dKr[:, q, :, 0]).max())
j = i[2]
for jj in (j - 2, j - 1, j, j + 1, j + 2):
    print("f", f[jj], "pair", out["1"][lev, 1, jj, 0], "generic", out["0"][lev, 1, jj, 0], "T: pair", out["1"][lev, 0, jj, 0], "generic", out["0"][lev, 0, jj, 0])
