#!/usr/bin/env python
"""BASELINE configs[4] (C5) as a batch: many nadir/slant paths x 1e4 frequencies with T + VMR Jacobians, driven the way
measurement_vecFromSensor drives the reference (src/m_rad.cc:321-343): concurrent host threads, each calling the
re-entrant C-ABI entry point with its own per-thread device workspace, one shared immutable catalog.

    python tests/tools/c5_batch.py --paths 64 --threads 4 > gpurun_out/c5_batch.json
"""
import argparse
import copy
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from arts_b200 import synth, wsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--paths", type=int, default=64)
ap.add_argument("--threads", type=int, default=4)
ap.add_argument("--lines", type=int, default=10_000)
ap.add_argument("--nf", type=int, default=10_000)
args = ap.parse_args()

base = synth.case_c5_single(n_lines=args.lines, nf=args.nf)
tg = (("T",), ("VMR", 0))
wsm.set_device(0)
cat = wsm.Catalog(base.cat)
zen = np.linspace(180.0, 120.0, args.paths)  # nadir ... 60 degrees off nadir
cases = []
for z in zen:
    c = copy.copy(base)
    c.r = base.r / abs(np.cos(np.deg2rad(z)))
    c.atm = copy.deepcopy(base.atm)
    c.atm.los = np.tile([z, 0.0], (base.np_, 1))
    cases.append(c)
out = [None] * args.paths
wsm.spectral_radClearskyEmission(cat, base.f, base.atm, base.r, base.I_bkg, jac_targets=tg, hse_derivative=1)  # warm


def run(nthreads, targets):
    nxt = [0]
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= args.paths:
                break
            c = cases[i]
            out[i] = wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=targets, hse_derivative=1)
        wsm.lib().ab200_release_thread_cache()

    th = [threading.Thread(target=work) for _ in range(nthreads)]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    return time.perf_counter() - t0


rep = {"paths": args.paths, "lines": args.lines, "nf": args.nf, "levels": base.np_, "targets": ["T", "VMR species 0"],
       "evals_per_path": float(args.lines) * args.nf * base.np_}
for name, targets in (("forward", ()), ("jacobian", tg)):
    for nt in (1, args.threads):
        s = run(nt, targets)
        rep[f"{name}_{nt}_threads"] = {"seconds": s, "paths_per_s": args.paths / s,
                                       "evals_per_s": rep["evals_per_path"] * args.paths / s}
# parity of one slant path against the oracle (sampled frequencies)
from tests import oracle_lib as orc  # noqa: E402

i = args.paths // 2
c = cases[i]
idx = np.unique(np.linspace(0, c.nf - 1, 24).astype(np.int64))
Ir, dIr = orc.clearsky_emission(c.cat, np.ascontiguousarray(c.f[idx]), c.atm, c.r, np.ascontiguousarray(c.I_bkg[idx]), targets=tg,
                                hse_derivative=1)
I, dI = out[i]
rep["check_path"] = i
rep["max_rel_dI"] = float((np.abs(I[idx, 0] - Ir[:, 0]) / Ir[:, 0]).max())
sc = np.abs(dIr[..., 0]).max(axis=(0, 1))
rep["max_rel_jac"] = float((np.abs(dI[idx][..., 0] - dIr[..., 0]).max(axis=(0, 1)) / sc).max())
rep["extrapolated_1e4_paths_s"] = 1e4 / rep[f"jacobian_{args.threads}_threads"]["paths_per_s"]
print(json.dumps(rep))
