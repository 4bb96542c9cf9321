#!/usr/bin/env python
"""Throughput of the AbsorptionBands XML loader (host code, one thread): a synthetic abs_bands file of --lines lines in
--bands bands, two broadeners with G0 / D0 models each, parsed straight into the SoA.

    python tools/xml_ingest_probe.py [--lines 1000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arts_b200 import _abi as abi  # noqa: E402
from arts_b200 import wsm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=1_000_000)
ap.add_argument("--bands", type=int, default=100)
args = ap.parse_args()
rng = np.random.default_rng(0)
per = args.lines // args.bands
parts = ['<?xml version="1.0"?>\n<arts format="ascii" version="1">\n<Map type="AbsorptionBand" key="QuantumIdentifier" nelem="%d">\n' % args.bands]
for b in range(args.bands):
    f0 = np.sort(rng.uniform(1e9, 3e12, per))
    a = rng.uniform(1e-12, 1e-3, per)
    e0 = rng.uniform(1e-22, 1e-20, per)
    g = rng.uniform(1e4, 3e4, (per, 2))
    parts.append('<QuantumIdentifier version="1"> H2O-161 v1 %d %d </QuantumIdentifier>\n' % (b, b))
    parts.append('<AbsorptionBand lineshape="VP_LTE" cutoff_type="None" cutoff_value="750000000000.0" nelem="%d">\n' % per)
    parts.extend("%r %r %r 9.0 11.0 0 0.0 0.0 296.0 2 Water 2 G0 T1 %r 0.75 D0 T5 -429.0 0.79 Bath 2 G0 T1 %r 0.7 D0 T0 120.5 2 J 3 2 Ka 1 2\n"
                 % (float(f0[i]), float(a[i]), float(e0[i]), float(g[i, 0]), float(g[i, 1])) for i in range(per))
    parts.append("</AbsorptionBand>\n")
parts.append("</Map>\n</arts>\n")
text = "".join(parts).encode()
t0 = time.perf_counter()
cat = wsm.abs_bandsReadXML(text=text, isotopologues=[("H2O-161", 0, 18.0106)], species_names={"Water": 0, "Bath": abi.SPECIES_BATH}, n_species=1)
dt = time.perf_counter() - t0
print(json.dumps({"lines": int(cat.n_lines), "bytes": len(text), "seconds": dt, "lines_per_s": cat.n_lines / dt, "MB_per_s": len(text) / dt / 1e6,
                  "note": "one host thread, includes the copy of the SoA into numpy arrays"}))
