#!/usr/bin/env python
"""Strong scaling of ONE host process over the GPUs of the box through ab200_multi_clearsky_emission (host buffers in, host
buffers out): BASELINE configs[3] (1e6 lines x --nf frequencies of the 1e6-point grid x 100 levels) with and without the
750 GHz ByLine cutoff, 1 device vs 2 / 4 / 8.  Checks that every multi-device result equals the one-device result bitwise.

    python tools/multi_probe.py [--nf 250000] [--cutoff-ghz 750] [--devices 1,2,4,8]
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
ap.add_argument("--nf", type=int, default=250_000)
ap.add_argument("--lines", type=int, default=1_000_000)
ap.add_argument("--levels", type=int, default=100)
ap.add_argument("--cutoff-ghz", type=float, default=750.0)
ap.add_argument("--devices", default="1,2,4,8")
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()

cut = args.cutoff_ghz * 1e9 if args.cutoff_ghz > 0 else None
case = synth.case_c4(n_lines=args.lines, nf=1_000_000, np_=args.levels, cutoff=cut)
stride = 1_000_000 // args.nf
f = np.ascontiguousarray(case.f[::stride][: args.nf])
bkg = np.ascontiguousarray(case.I_bkg[::stride][: args.nf])
navail = wsm.device_count()
rep = {"workload": f"configs[3]: {args.lines} lines x {args.nf} frequencies (every {stride}-th point of the 1e6-point grid) x {args.levels} levels, "
                   + (f"ByLine cutoff {args.cutoff_ghz:g} GHz" if cut else "no cutoff") + ", linsrc; one host process, host buffers",
       "devices_visible": navail, "runs": []}
ref = None
for n in [int(x) for x in args.devices.split(",")]:
    if n > navail:
        continue
    m = wsm.MultiDevice(case.cat, n_devices=n)
    I, _ = wsm.spectral_radClearskyEmission(m, f, case.atm, case.r, bkg)  # builds the workspaces
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        I, _ = wsm.spectral_radClearskyEmission(m, f, case.atm, case.r, bkg)
        ts.append(time.perf_counter() - t0)
    if ref is None:
        ref = I
    t = min(ts)
    rep["runs"].append({"devices": n, "seconds": t, "evals_per_s": float(args.lines) * args.nf * args.levels / t,
                        "bitwise_equal_to_first": bool(np.array_equal(I, ref))})
    m.close()
t1 = rep["runs"][0]["seconds"] * rep["runs"][0]["devices"]
for r in rep["runs"]:
    r["speedup_vs_1_device"] = t1 / r["seconds"] if rep["runs"][0]["devices"] == 1 else None
print(json.dumps(rep))
