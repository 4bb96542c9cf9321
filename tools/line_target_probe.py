#!/usr/bin/env python
"""Cost of line-parameter Jacobian rows on one configs[4] path (1e4 lines x 1e4 frequencies x 100 levels): a pass of four
line targets only visits the tiles that hold those lines, so it should cost a small fraction of a temperature row.

    python tools/line_target_probe.py [--reps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import _abi as abi  # noqa: E402
from arts_b200 import synth, wsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
c = synth.case_c5_single()
wsm.set_device(0)
cat = wsm.Catalog(c.cat)
lines = [123, 4567, 7001, 9990]
sets = {
    "T": (("T",),),
    "4 line targets (f0, a, G0.X0, D0.X0 of four lines)": (("line_f0", lines[0]), ("line_a", lines[1]),
                                                         ("line_ls", lines[2], abi.VAR_G0, abi.SPECIES_BATH, 0),
                                                         ("line_ls", lines[3], abi.VAR_D0, abi.SPECIES_BATH, 0)),
    "8 line targets": tuple(("line_f0", l) for l in lines) + tuple(("line_e0", l) for l in lines),
}
rep = {"workload": f"C5 one path: {c.cat.n_lines} lines x {c.nf} frequencies x {c.np_} levels"}
for name, tg in sets.items():
    p = wsm.Path(cat, c.nf, c.np_, len(tg))
    p.upload(c.f, c.atm, c.r, c.I_bkg, targets=tg)
    p.run_propmat(); p.sync()
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        p.run_propmat(); p.sync()
        ts.append(time.perf_counter() - t0)
    rep[name] = {"propmat_ms": 1e3 * float(np.median(ts))}
    p.close()
print(json.dumps(rep))
