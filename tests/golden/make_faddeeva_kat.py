"""Extracts the reference's own w(z) known-answer test (57 points, WolframAlpha/Maple values)
from /root/reference/3rdparty/Faddeeva/Faddeeva.cc:4041-4195 into tests/golden/faddeeva_kat.json.

Run in the CPU container (the GPU box has no /root/reference); the JSON is committed.
Numbers are kept as the source's decimal strings (45+ digits) and parsed with float().
"""
import json
import os
import re
import sys

REF = os.environ.get("ARTS_REFERENCE", "/root/reference")
src = open(os.path.join(REF, "3rdparty/Faddeeva/Faddeeva.cc")).read()
test = src[src.index("w(z) tests"):]


def block(name):
    m = re.search(r"cmplx %s\[NTST\] = \{(.*?)\};" % name, test, re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    body = re.sub(r"//[^\n]*", "", body)
    pts = re.findall(r"C\(\s*([^,()]+?)\s*,\s*([^,()]+?)\s*\)", body)
    return [[a.strip(), b.strip()] for a, b in pts]


z, w = block("z"), block("w")
assert len(z) == 57 and len(w) == 57, (len(z), len(w))
out = {"source": "3rdparty/Faddeeva/Faddeeva.cc:4041-4195 (TEST_FADDEEVA, w(z) suite)", "threshold": 1e-13,
       "z": z, "w": w}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "faddeeva_kat.json")
json.dump(out, open(path, "w"), indent=0)
print("wrote", path, len(z), "points")
