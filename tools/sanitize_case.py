#!/usr/bin/env python
"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck, one tool per gpurun call):
scalar + Zeeman forward, ByLine cutoff, Jacobians, linprop, wind, un-fused entry points."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import synth, wsm
wsm.set_device(0)
tg = (("T",), ("VMR", 0))
c = synth.tiny_case(nl=300, nf=700, np_=6)
wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1)
wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option="linprop")
c.atm.wind = np.full((c.np_, 3), 20.0)
wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option="constant")
c = synth.case_c3(nf=38 * 8, np_=5, los=(120.0, 30.0))
I, dI, K = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, return_propmat=True)
c = synth.case_c1(nl=300, nf=600, cutoff=2e9)
K1, dK1 = wsm.spectral_propmat_pathFromPath(c.cat, c.f, c.atm, jac_targets=tg)
tm = wsm.spectral_tramat_pathFromPath(K, None, np.full(4, 500.0), np.linspace(200, 250, 5), "linsrc")
J, dJ = wsm.spectral_rad_srcvec_pathFromPropmat(K, synth.case_c3(nf=38 * 8, np_=5).f, np.linspace(200, 250, 5))
wsm.spectral_radStepByStepEmission(tm, J, dJ, np.zeros((K.shape[1], 4)))
wsm.spectral_radApplyPlanckTb(I, synth.case_c3(nf=38 * 8, np_=5).f)
print("sanitize pass done", float(np.abs(I).sum()), float(np.abs(K1).sum()))
