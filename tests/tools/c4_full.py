#!/usr/bin/env python
"""BASELINE configs[3] at FULL size on one B200: 1e6 lines x 1e6 frequencies x 100 levels = 1e14
(line, frequency, level) evaluations + 1e8 Stokes steps.  Prints one JSON report (profiles/).

    python tests/tools/c4_full.py [--cutoff-ghz 750]   # (ii) the realistic ByLine-cutoff variant
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from arts_b200 import roofline, synth, wsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cutoff-ghz", type=float, default=None)
ap.add_argument("--nf", type=int, default=1_000_000)
ap.add_argument("--check", type=int, default=16, help="frequencies re-computed alone and by the CPU oracle")
args = ap.parse_args()

t0 = time.time()
c = synth.case_c4(nf=args.nf, cutoff=None if args.cutoff_ghz is None else args.cutoff_ghz * 1e9)
wsm.set_device(0)
stream = torch.cuda.current_stream()
cat = wsm.Catalog(c.cat)
path = wsm.Path(cat, c.nf, c.np_, stream=stream.cuda_stream)
path.upload(c.f, c.atm, c.r, c.I_bkg)
setup_s = time.time() - t0
hist = path.region_histogram(100_000, seed=4)
fl, regions = roofline.flops_per_eval(hist)
dfma, _ = wsm.measure_dfma_peak(20000)
path.set_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
path.run_propmat()
path.run_stokes()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
kt = path.timings()
I = np.empty((c.nf, 4))
path.download(I=I)
evals = float(c.n_lines) * c.nf * c.np_
evaluated = hist[:5].sum() / max(hist[7], 1)
rep = {
    "workload": f"C4: {c.n_lines} lines x {c.nf} frequencies x {c.np_} levels" + (f", ByLine cutoff {args.cutoff_ghz} GHz" if args.cutoff_ghz else ", no cutoff"),
    "ms": ms, "evals_per_s_nominal": evals / (ms * 1e-3), "fraction_of_pairs_inside_cutoff": evaluated,
    "kernel_ms": {k: v[0] for k, v in kt.items()}, "kernel_launches": {k: v[1] for k, v in kt.items()},
    "flop_per_eval": fl, "regions": regions, "dfma_peak_tflops": dfma,
    "roofline_frac_sum": fl * evals * evaluated / (sum(kt[k][0] for k in ("sum_real", "sum_cplx")) * 1e-3) / 1e12 / dfma,
    "stokes_steps_per_s": float(c.nf) * c.np_ / (kt["stokes"][0] * 1e-3), "setup_s": setup_s,
    "tb_min_max": [float(wsm.spectral_radApplyPlanckTb(I[::1000], c.f[::1000])[:, 0].min()),
                   float(wsm.spectral_radApplyPlanckTb(I[::1000], c.f[::1000])[:, 0].max())],
}
if args.check:
    from tests import oracle_lib as orc

    idx = np.unique(np.linspace(0, c.nf - 1, args.check).astype(np.int64))
    fs, bs = np.ascontiguousarray(c.f[idx]), np.ascontiguousarray(c.I_bkg[idx])
    p2 = wsm.Path(cat, len(idx), c.np_)
    p2.set_grid_bounds(np.tile([c.f[0], c.f[-1]], (c.np_, 1)))
    p2.upload(fs, c.atm, c.r, bs)
    p2.run_propmat()
    p2.run_stokes()
    Is = np.empty((len(idx), 4))
    p2.download(I=Is)
    rep["sample_bitwise_equal"] = bool(np.array_equal(Is, I[idx]))
    t1 = time.time()
    Ir, _ = orc.clearsky_emission(c.cat, fs, c.atm, c.r, bs)
    rep["oracle_s"] = time.time() - t1
    rep["oracle_evals_per_s"] = float(c.n_lines) * len(idx) * c.np_ * evaluated / rep["oracle_s"]
    rep["oracle_threads"] = orc.num_threads()
    tb, tbr = wsm.spectral_radApplyPlanckTb(Is, fs), orc.planck_tb(fs, Ir)
    rep["max_abs_dTb_K"] = float(np.abs(tb - tbr).max())
    rep["max_rel_dI"] = float((np.abs(Is[:, 0] - Ir[:, 0]) / np.abs(Ir[:, 0])).max())
print(json.dumps(rep))
