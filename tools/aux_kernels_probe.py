#!/usr/bin/env python
"""Times the HBM-bound kernels around the line sum on the bench shape (configs[1]: 100 levels x 1e5 frequencies): the other
absorption terms added into the resident K (standard continua, CIA, lookup-table extraction) and the fused Stokes chain, against
the bytes each has to move (K.A read + written: 16 B per (frequency, level); the chain: 56 B read).

    python tools/aux_kernels_probe.py > gpurun_out/aux_kernels.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import _abi as abi  # noqa: E402
from arts_b200 import synth, wsm  # noqa: E402

c = synth.case_c2(lines_per_species=200)  # the line sum is not what is timed here
nf, np_ = c.nf, c.np_
wsm.set_device(0)
cat = wsm.Catalog(c.cat)
p = wsm.Path(cat, nf, np_, 0)
p.upload(c.f, c.atm, c.r, c.I_bkg)
p.run_propmat(); p.sync()


def timed(fn, reps=20):
    fn(); p.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    p.sync()
    return (time.perf_counter() - t0) / reps


rng = np.random.default_rng(1)
fg = np.linspace(c.f[0], c.f[-1], 400)
Tg = np.array([150.0, 200.0, 250.0, 300.0, 350.0])
cia = wsm.Cia([abi.CiaRecord(3, 4, [(fg, Tg, 1e-56 * (1 + rng.random((400, 5))))]), abi.CiaRecord(3, 3, [(fg, Tg, 1e-56 * (1 + rng.random((400, 5))))])])
ref = c.atm if c.atm.P[0] > c.atm.P[-1] else c.atm.reversed()
ref = ref.take(np.arange(np_))  # the path levels themselves: offsets and ratios are exactly 0 and 1
fl = np.linspace(c.f[0], c.f[-1], 2000)
tables = [abi.LookupTable(species=s, f_grid=fl, log_p_grid=np.log(ref.P), t_atmref=np.array(ref.T), xsec=1e-26 * (1 + rng.random((5, 5 if s == 0 else 1, ref.np_, 2000))),
                          t_pert=np.linspace(-40, 40, 5), w_pert=np.geomspace(0.05, 20, 5) if s == 0 else None,
                          water_atmref=np.array(ref.vmr[:, 0]) if s == 0 else None) for s in range(2)]
lut = wsm.Lookup(tables)
species = {"H2O": 0, "O2": 3, "N2": 4}
models = ["O2-SelfContStandardType", "N2-SelfContStandardType", "H2O-ForeignContStandardType", "H2O-SelfContStandardType"]
elems = float(nf) * np_
rep = {"workload": f"{np_} levels x {nf} frequencies", "peak_hbm_gbs": 6538.3, "kernels": {}}
for name, fn, bytes_per in (
        ("standard continua (4 models)", lambda: p.add_predefined(models, species), 16.0),
        ("CIA (2 pairs, 400 x 5 tables)", lambda: p.add_cia(cia), 16.0),
        ("lookup extraction (2 tables, orders p5 t4 w4 f1)", lambda: p.add_lookup(lut, h2o_species=0, p_interp_order=5, t_interp_order=4, water_interp_order=4, f_interp_order=1, zero_init=False), 16.0),
        ("fused Stokes chain (linsrc, scalar)", lambda: p.run_stokes(), 56.0 + 32.0 / np_)):
    t = timed(fn)
    rep["kernels"][name] = {"ms": 1e3 * t, "algorithmic_GBps": elems * bytes_per / t / 1e9, "frac_of_hbm_peak": elems * bytes_per / t / 1e9 / 6538.3}
print(json.dumps(rep))
