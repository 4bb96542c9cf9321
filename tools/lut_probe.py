#!/usr/bin/env python
"""Absorption lookup tables on the GPU: (1) abs_lookup_dataPrecompute with the line-by-line sum (the table constructor of
src/core/lookup/lookup_map.cpp:22-131 is the hot path over nt * nw * np perturbed levels), (2) extraction on a 100-level path
(spectral_propmatAddLookup, src/m_lookup.cc) against the line-by-line run of the same path.

    python tools/lut_probe.py > gpurun_out/lut_probe.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import synth, wsm  # noqa: E402

nf = 20_000
c = synth.case_c2(nf=nf, np_=100)
full = c.atm if c.atm.P[0] > c.atm.P[-1] else c.atm.reversed()  # lookup-table profiles run surface -> top
ref = full.take(np.unique(np.r_[np.arange(0, 100, 2), 99]))        # every other level of the path's range: 51 pressures
np_ref = ref.np_
wsm.set_device(0)
cat = wsm.Catalog(c.cat)
t_pert, w_pert = np.linspace(-30, 30, 5), np.geomspace(0.1, 10, 5)
wsm.abs_lookup_dataPrecompute(cat, ref, c.f[:256], 1)  # warm-up
t0 = time.perf_counter()
tables = [wsm.abs_lookup_dataPrecompute(cat, ref, c.f, s, temperature_perturbation=t_pert,
                                        water_perturbation=w_pert if s == 0 else None, h2o_species=0) for s in range(c.cat.n_species)]
t_pre = time.perf_counter() - t0
lines_per_species = len(c.cat.f0) // c.cat.n_species
evals = float(lines_per_species) * nf * np_ref * len(t_pert) * (len(w_pert) + c.cat.n_species - 1)
lut = wsm.Lookup(tables)
path_case = c
p = wsm.Path(cat, nf, 100, 0)
p.upload(path_case.f, path_case.atm, path_case.r, path_case.I_bkg)
p.set_timing(True)
p.run_propmat(); p.run_stokes(); p.sync()
I = np.empty((nf, 4)); p.download(I=I)
ts = []
for _ in range(5):
    p.sync(); t0 = time.perf_counter()
    p.run_propmat(); p.sync()
    ts.append(time.perf_counter() - t0)
t_lbl = float(np.median(ts))
kw = dict(h2o_species=0, p_interp_order=5, t_interp_order=4, water_interp_order=4, f_interp_order=0)
p.add_lookup(lut, **kw); p.sync()
ts = []
for _ in range(5):
    p.sync(); t0 = time.perf_counter()
    p.add_lookup(lut, **kw); p.sync()
    ts.append(time.perf_counter() - t0)
t_lut = float(np.median(ts))
p.run_stokes()
Il = np.empty((nf, 4)); p.download(I=Il)
tb, tbl = wsm.spectral_radApplyPlanckTb(I, path_case.f)[:, 0], wsm.spectral_radApplyPlanckTb(Il, path_case.f)[:, 0]
size_mb = sum(t.xsec.nbytes for t in tables) / 1e6
print(json.dumps({
    "workload": f"C2 catalog ({len(c.cat.f0)} lines, 5 species), {nf} frequencies; table: {np_ref} pressures x {len(t_pert)} temperature offsets"
                f" x {len(w_pert)} water ratios (H2O only); path: 100 levels",
    "precompute": {"seconds": t_pre, "evals": evals, "evals_per_s": evals / t_pre, "table_MB": size_mb},
    "extraction_orders_p_t_w_f": [5, 4, 4, 0],
    "path_100_levels": {"line_by_line_ms": 1e3 * t_lbl, "lookup_ms": 1e3 * t_lut, "speedup": t_lbl / t_lut,
                        "max_abs_dTb_K": float(np.abs(tb - tbl).max())}}))
