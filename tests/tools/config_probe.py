#!/usr/bin/env python
"""Device-timed throughput of the BASELINE configs that are not the bench workload (C1, C3, C5 one path),
with a parity spot check against the CPU oracle on a frequency sample.  One JSON object on stdout.

    python tests/tools/config_probe.py > gpurun_out/configs.json
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from arts_b200 import roofline, synth, wsm  # noqa: E402
from tests import oracle_lib as orc  # noqa: E402

wsm.set_device(0)
stream = torch.cuda.current_stream()
dfma, _ = wsm.measure_dfma_peak(20000)
out = {"dfma_peak_tflops": dfma}


def probe(name, c, targets=(), n_check=32, reps=5):
    cat = wsm.Catalog(c.cat)
    nq = len(targets)
    p = wsm.Path(cat, c.nf, c.np_, nq, stream=stream.cuda_stream)
    p.upload(c.f, c.atm, c.r, c.I_bkg, rte_option=c.rte_option, targets=targets, hse_derivative=1 if nq else 0)
    counts = cat.counts()
    nsub = float(sum(counts))
    for _ in range(2):
        p.run_propmat()
        if c.np_ > 1:
            p.run_stokes()
    torch.cuda.synchronize()
    p.timings()
    p.set_timing(True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(reps):
        p.run_propmat()
    e[1].record()
    for _ in range(reps):
        if c.np_ > 1:
            p.run_stokes()
    e[2].record()
    torch.cuda.synchronize()
    kt = p.timings()
    ms_pm, ms_st = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps
    hist = p.region_histogram(200_000, seed=3)
    fl, regions = roofline.flops_per_eval(hist)
    evals = nsub * c.nf * c.np_
    rep = {"sub_lines": nsub, "nf": c.nf, "levels": c.np_, "nq": nq, "evals": evals, "propmat_ms": ms_pm, "stokes_ms": ms_st,
           "evals_per_s": evals / (ms_pm * 1e-3), "stokes_steps_per_s": (c.nf * c.np_ / (ms_st * 1e-3)) if c.np_ > 1 and ms_st > 0 else None,
           "kernel_ms": {k: v[0] / max(v[1], 1) for k, v in kt.items()}, "flop_per_eval": fl, "regions": regions,
           "fp64_algorithmic_tflops_forward": fl * evals / (sum(kt[k][0] for k in ("sum_real", "sum_cplx")) / reps * 1e-3) / 1e12}
    # parity spot check
    idx = np.unique(np.linspace(0, c.nf - 1, n_check).astype(np.int64))
    fs, bs = np.ascontiguousarray(c.f[idx]), np.ascontiguousarray(c.I_bkg[idx])
    if c.np_ > 1:
        I = np.empty((c.nf, 4))
        dI = np.empty((c.nf, c.np_, nq, 4)) if nq else None
        p.download(I=I, dI=dI)
        Ir, dIr = orc.clearsky_emission(c.cat, fs, c.atm, c.r, bs, rte_option=c.rte_option, targets=targets, hse_derivative=1 if nq else 0)
        tb, tbr = wsm.spectral_radApplyPlanckTb(I[idx], fs), orc.planck_tb(fs, Ir)
        rep["max_abs_dTb_K"] = float(np.abs(tb - tbr).max())
        if nq:
            sc = np.abs(dIr[..., 0]).max(axis=(0, 1))
            rep["max_rel_dI_jac"] = float((np.abs(dI[idx][..., 0] - dIr[..., 0]).max(axis=(0, 1)) / sc).max())
    else:
        K = np.empty((1, c.nf, 7))
        p.download(K=K)
        Kr, _ = orc.propmat_levels(c.cat, fs, c.atm)
        rep["max_rel_dK"] = float((np.abs(K[0, idx, 0] - Kr[0, :, 0]) / Kr[0, :, 0]).max())
    out[name] = rep
    p.close()
    cat.close()


probe("C1 (1k lines x 1e4 freqs x 1 level, propmat only)", synth.case_c1(), reps=50)
probe("C3 (O2 Zeeman, 4332 sub-lines x 1e5 freqs x 50 levels, polarised linsrc chain)", synth.case_c3())
probe("C3 constant", synth.case_c3(rte_option="constant"))
probe("C5 one path forward (1e4 lines x 1e4 freqs x 100 levels)", synth.case_c5_single(), reps=5)
probe("C5 one path, T + VMR Jacobians", synth.case_c5_single(), targets=(("T",), ("VMR", 0)), reps=3)
print(json.dumps(out))
