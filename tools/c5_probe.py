#!/usr/bin/env python
"""Times one path of BASELINE configs[4] (C5): 1e4 lines x 1e4 frequencies x 100 levels with T + VMR Jacobians."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arts_b200 import synth, wsm
c = synth.case_c5_single()
tg = (("T",), ("VMR", 0))
wsm.set_device(0)
cat = wsm.Catalog(c.cat)
stream = torch.cuda.current_stream()
wsm.set_thread_stream(stream.cuda_stream)
rep = {}
for name, targets in (("forward", ()), ("T+VMR jacobians", tg)):
    wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=targets)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        wsm.spectral_radClearskyEmission(cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=targets)
    torch.cuda.synchronize()
    rep[name + " ms/path (e2e, host buffers)"] = 1e3 * (time.perf_counter() - t0) / n
p = wsm.Path(cat, c.nf, c.np_, 2, stream=stream.cuda_stream)
p.upload(c.f, c.atm, c.r, c.I_bkg, targets=tg)
p.set_timing(True)
for _ in range(3):
    p.run_propmat(); p.run_stokes()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); p.run_propmat(); e1.record(); torch.cuda.synchronize()
rep["propmat+jac ms (resident)"] = e0.elapsed_time(e1)
e0.record(); p.run_stokes(); e1.record(); torch.cuda.synchronize()
rep["stokes+jac ms (resident)"] = e0.elapsed_time(e1)
rep["evals"] = float(c.n_lines) * c.nf * c.np_
print(json.dumps(rep))
