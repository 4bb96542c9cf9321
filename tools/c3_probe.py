#!/usr/bin/env python
"""BASELINE configs[2] (C3): O2 60-GHz band with Zeeman splitting, 4332 sub-lines x 1e5 frequencies x 50 levels, full
polarised propagation matrix + polarised linsrc Stokes chain.  Device-resident timing of both stages; the ncu target
for lbl_sum_cplx_kernel (tools/ncu_cplx.sh).

    python tools/c3_probe.py [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import synth, wsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
c = synth.case_c3()
wsm.set_device(0)
cat = wsm.Catalog(c.cat)
p = wsm.Path(cat, c.nf, c.np_, 0)
p.upload(c.f, c.atm, c.r, c.I_bkg, rte_option=c.rte_option)
p.set_timing(True)
p.run_propmat(); p.run_stokes(); p.sync(); p.timings()
ts = []
for _ in range(args.reps):
    t0 = time.perf_counter()
    p.run_propmat(); p.sync()
    ts.append(time.perf_counter() - t0)
    p.run_stokes(); p.sync()
tm = p.timings()
n_sub = sum(cat.counts())
evals = float(n_sub) * c.nf * c.np_
rep = {"workload": f"C3: {n_sub} Zeeman sub-lines x {c.nf} frequencies x {c.np_} levels", "evals": evals,
       "propmat_ms": 1e3 * float(np.median(ts)), "evals_per_s": evals / float(np.median(ts)),
       "kernel_ms": {k: v[0] / max(v[1], 1) for k, v in tm.items()}, "regions": p.region_histogram().tolist()}
print(json.dumps(rep))
# Jacobian rows of the Zeeman configuration: temperature + the three magnetic-field components, and the three wind rows
p.close() if hasattr(p, "close") else None
for name, tg in (("T+mag_uvw", (("T",), ("mag_u",), ("mag_v",), ("mag_w",))), ("wind_uvw", (("wind_u",), ("wind_v",), ("wind_w",)))):
    pj = wsm.Path(cat, c.nf, c.np_, len(tg))
    pj.upload(c.f, c.atm, c.r, c.I_bkg, rte_option=c.rte_option, targets=tg, hse_derivative=1)
    pj.run_propmat(); pj.run_stokes(); pj.sync()
    tp, tsk = [], []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        pj.run_propmat(); pj.sync()
        t1 = time.perf_counter()
        pj.run_stokes(); pj.sync()
        tp.append(t1 - t0); tsk.append(time.perf_counter() - t1)
    print(json.dumps({"targets": name, "propmat_ms": 1e3 * float(np.median(tp)), "stokes_jac_ms": 1e3 * float(np.median(tsk))}))
    pj.close() if hasattr(pj, "close") else None
