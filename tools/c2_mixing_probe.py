#!/usr/bin/env python
"""BASELINE configs[1] with first-order line mixing on every line (Y = T1 model): all bands are complex segments with
pol = no, summed by lbl_sum_cplx_kernel's real-only far loop.  Device-timed propagation-matrix stage."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import _abi as abi  # noqa: E402
from arts_b200 import synth, wsm  # noqa: E402

c = synth.case_c2()
n = len(c.cat.ls_species)
c.cat.ls_type[:, abi.VAR_Y] = abi.TM_T1
c.cat.ls_X[:, abi.VAR_Y, 0] = np.random.default_rng(1).uniform(-1e-7, 1e-7, n)
c.cat.ls_X[:, abi.VAR_Y, 1] = 0.8
wsm.set_device(0)
cat = wsm.Catalog(c.cat)
p = wsm.Path(cat, c.nf, c.np_, 0)
p.upload(c.f, c.atm, c.r, c.I_bkg)
p.set_timing(True)
p.run_propmat(); p.sync(); p.timings()
for _ in range(3):
    p.run_propmat()
p.sync()
tm = p.timings()
ms = tm["sum_cplx"][0] / tm["sum_cplx"][1] * (tm["sum_cplx"][1] / 3)
evals = float(len(c.cat.f0)) * c.nf * c.np_
print(json.dumps({"workload": "C2 with line mixing on every line (complex segments, pol = no)", "sum_cplx_ms_per_step": ms,
                  "evals_per_s": evals / (ms * 1e-3), "kernel_ms": {k: v[0] / 3 for k, v in tm.items()}}))
