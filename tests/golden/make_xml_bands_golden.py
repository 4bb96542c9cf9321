#!/usr/bin/env python
"""Golden for tests/test_xml_bands.py from the reference's own fixture tests/core/nlte/nlte_lines.xml: the first three of
its nine bands as a (valid, smaller) abs_bands file plus a few numbers read off the text by hand-written slicing, so that
the test can run without /root/reference.

    python tests/golden/make_xml_bands_golden.py   (in the build container, where /root/reference exists)
"""
import json
import os
import re

src = open("/root/reference/tests/core/nlte/nlte_lines.xml").read()
pairs = re.findall(r'(<QuantumIdentifier.*?</AbsorptionBand>\n)', src, re.S)
keep = pairs[:3]
text = '<?xml version="1.0"?>\n<arts format="ascii" version="1">\n<Map type="AbsorptionBand" key="QuantumIdentifier" nelem="3">\n' + \
       "".join(keep) + "</Map>\n</arts>\n"
first = keep[0].split("\n")[2].split()
out = {
    "source": "tests/core/nlte/nlte_lines.xml, bands 0-2 of %d" % len(pairs),
    "text": text,
    "n_bands_in_fixture": len(pairs),
    "band0": {"isot": keep[0].split(">")[1].split()[0], "f0": float(first[0]), "a": float(first[1]), "e0": float(first[2]),
              "gu": float(first[3]), "gl": float(first[4]), "T0": float(first[8]), "n_broadeners": int(first[9]),
              "first_broadener": first[10], "G0_type": first[13], "G0_X0": float(first[14]), "G0_X1": float(first[15]),
              "D0_type": first[17], "D0_X0": float(first[18])},
}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "xml_bands_fixture.json"), "w"), indent=1)
print("bands kept:", len(keep), "first line tokens:", len(first))
