#!/usr/bin/env python
"""One path of BASELINE configs[4] (1e4 lines x 1e4 frequencies x 100 levels) with T + VMR Jacobians: wall time of
the propagation-matrix stage with and without targets (device-resident inputs, stream synchronised on both sides).
Used as the ncu target for lbl_sum_jac_kernel (tools/ncu_jac.sh).

    python tools/c5_jac_probe.py [--reps 5]
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
ap.add_argument("--lines", type=int, default=10_000)
ap.add_argument("--nf", type=int, default=10_000)
args = ap.parse_args()
c = synth.case_c5_single(n_lines=args.lines, nf=args.nf)
wsm.set_device(0)
cat = wsm.Catalog(c.cat)
rep = {"workload": f"C5 one path: {args.lines} lines x {args.nf} frequencies x {c.np_} levels"}
for name, tg in (("forward", ()), ("T+VMR", (("T",), ("VMR", 0))), ("T+3VMR", (("T",), ("VMR", 0), ("VMR", 1), ("VMR", 3)))):
    p = wsm.Path(cat, c.nf, c.np_, len(tg))
    p.upload(c.f, c.atm, c.r, c.I_bkg, targets=tg, hse_derivative=1)
    p.run_propmat(); p.sync()
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        p.run_propmat(); p.sync()
        ts.append(time.perf_counter() - t0)
    t1 = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        p.run_stokes(); p.sync()
        t1.append(time.perf_counter() - t0)
    rep[name] = {"propmat_ms": 1e3 * float(np.median(ts)), "stokes_ms": 1e3 * float(np.median(t1))}
    p.close()
print(json.dumps(rep))
