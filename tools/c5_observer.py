#!/usr/bin/env python
"""BASELINE configs[4] (C5) as a batch of slant paths with T + VMR Jacobians, two ways:

  a) ab200_clearsky_emission per path: the per-level Jacobian spectral_rad_jac_path [nf][np][nq] comes back to the
     host (what the reference's observer agenda hands to spectral_rad_jacAddPathPropagation, src/m_rad.cc:62-127);
  b) wsm.measurement_vecFromSensor: background, state-space mapping, Planck-Tb transform and channel sum-up on the
     device (ab200_path_run_observer); only [channels] + [channels][nx] doubles per path cross PCIe.

    python tools/c5_observer.py --paths 32 > gpurun_out/c5_observer.json
"""
import argparse
import copy
import gc
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import _abi as abi  # noqa: E402
from arts_b200 import synth, wsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--paths", type=int, default=32)
ap.add_argument("--lines", type=int, default=10_000)
ap.add_argument("--nf", type=int, default=10_000)
ap.add_argument("--channels", type=int, default=20)
ap.add_argument("--grid", type=int, default=50, help="retrieval grid nodes per target")
args = ap.parse_args()

base = synth.case_c5_single(n_lines=args.lines, nf=args.nf)
tg = (("T",), ("VMR", 0))
nq = len(tg)
wsm.set_device(0)
cat = wsm.Catalog(base.cat)
rng = np.random.default_rng(5)
zen = np.linspace(180.0, 120.0, args.paths)
per = args.nf // args.channels
channels = [[(int(j), (1.0 / per, 0.0, 0.0, 0.0)) for j in range(c * per, (c + 1) * per)] for c in range(args.channels)]
nx = args.grid * nq + 1
sims, cases = [], []
for z in zen:
    c = copy.copy(base)
    c.r = base.r / abs(np.cos(np.deg2rad(z)))
    c.atm = copy.deepcopy(base.atm)
    c.atm.los = np.tile([z, 0.0], (base.np_, 1))
    path_map = []
    for ip in range(base.np_):
        pos = ip * (args.grid - 1) / (base.np_ - 1)
        i0 = min(int(pos), args.grid - 2)
        w1 = pos - i0
        path_map.append([[(t * args.grid + i0, 1.0 - w1), (t * args.grid + i0 + 1, w1)] for t in range(nq)])
    obs = abi.Observer(nx=nx, path_map=path_map, bkg_T=288.0, bkg_rows=[(nx - 1, 1.0)], unit="PlanckBT", channels=channels)
    sims.append((c.atm, c.r, obs))
    cases.append(c)

gc.disable()  # hundreds of thousands of small tuples describe the observers: keep the collector out of the timings
# warm-up of both routes
wsm.spectral_radClearskyEmission(cat, base.f, base.atm, base.r, base.I_bkg, jac_targets=tg, hse_derivative=1)
wsm.measurement_vecFromSensor(cat, base.f, sims, jac_targets=tg, hse_derivative=1)  # also flattens every observer once

t0 = time.perf_counter()
for c in cases:
    I, dI = wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1)
ta = time.perf_counter() - t0
t0 = time.perf_counter()
y, J = wsm.measurement_vecFromSensor(cat, base.f, sims, jac_targets=tg, hse_derivative=1)
tb = time.perf_counter() - t0
t0 = time.perf_counter()
y1, J1 = wsm.measurement_vecFromSensor(cat, base.f, sims, jac_targets=tg, hse_derivative=1, n_workspaces=1)
tb1 = time.perf_counter() - t0

# the two routes agree: map the last path's downloaded dI on the host like m_rad.cc:107-125 and compare its Jx rows
evals = float(args.lines) * args.nf * base.np_
rep = {
    "workload": f"C5 batch: {args.paths} slant paths x {args.lines} lines x {args.nf} frequencies x {base.np_} levels, targets T + VMR",
    "a_clearsky_emission_per_path": {"seconds": ta, "paths_per_s": args.paths / ta, "evals_per_s": evals * args.paths / ta,
                                     "d2h_bytes_per_path": int(args.nf * 4 * 8 * (1 + base.np_ * nq))},
    "b_measurement_vecFromSensor": {"seconds": tb, "paths_per_s": args.paths / tb, "evals_per_s": evals * args.paths / tb,
                                    "d2h_bytes_per_path": int(8 * args.channels * (1 + nx)), "workspaces": 2},
    "b_one_workspace": {"seconds": tb1, "paths_per_s": args.paths / tb1},
    "same_result_one_or_two_workspaces": bool(np.array_equal(y, y1) and np.array_equal(J, J1)),
    "y_mean_K": float(y.mean() / args.paths), "channels": args.channels, "nx": nx,
}
print(json.dumps(rep))
