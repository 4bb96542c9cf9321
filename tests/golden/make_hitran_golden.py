#!/usr/bin/env python
"""Golden vector of the HITRAN reader: the reference's own fixture tests/hitran/single_line.par (one H2O record, read by
tests/hitran/read.py).  Its first 160 columns are the `par` part; the quantum-number tail belongs to the statep / statepp
formatters, which are outside this path.  Run in the CPU container (needs /root/reference):

    python tests/golden/make_hitran_golden.py
"""
import json
import os

here = os.path.dirname(os.path.abspath(__file__))
ref = os.environ.get("ARTS_REFERENCE", "/root/reference")
rec = open(os.path.join(ref, "tests", "hitran", "single_line.par")).readline().rstrip("\n")[:160]
assert len(rec) == 160
# the columns, read off the record by eye (HITRAN 2004+ 160-character format)
expected = {"M": 1, "I": "1", "nu_cm-1": 0.072049, "A_s-1": 4.668e-12, "gamma_air_cm-1_atm-1": 0.0946,
            "gamma_self_cm-1_atm-1": 0.391, "E_cm-1": 1922.8289, "n_air": 0.73, "delta_cm-1_atm-1": 0.002760,
            "g_upp": 9.0, "g_low": 11.0}
json.dump({"record": rec, "columns": expected}, open(os.path.join(here, "hitran_single_line.json"), "w"), indent=1)
print(rec)
