#!/usr/bin/env python
"""Determinism check at the bench shape: repeated steps on the resident path must give bit-identical spectra."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arts_b200 import synth, wsm
wsm.set_device(0)
c = synth.case_c2()
cat = wsm.Catalog(c.cat)
stream = torch.cuda.current_stream()
p = wsm.Path(cat, c.nf, c.np_, stream=stream.cuda_stream)
p.upload(c.f, c.atm, c.r, c.I_bkg, rte_option="linsrc")
def get():
    I = np.empty((c.nf, 4)); p.download(I=I); return I
p.run_propmat(); p.sync(); p.run_stokes(); ref = get()
for mode in ("propmat,stokes back to back", "propmat, sync, stokes", "stokes only"):
    for rep in range(4):
        if mode.startswith("propmat,stokes"):
            p.run_propmat(); p.run_stokes()
        elif mode.startswith("propmat, sync"):
            p.run_propmat(); p.sync(); p.run_stokes()
        else:
            p.run_stokes()
        I = get()
        d = np.abs(I[:, 0] - ref[:, 0]) / ref[:, 0]
        bad = np.nonzero(d > 0)[0]
        print(mode, rep, "n diff", len(bad), "max rel", float(d.max()), "bad idx", bad[:6].tolist())
