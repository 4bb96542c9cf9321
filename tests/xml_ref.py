"""Pure-Python restatement of the reference's AbsorptionBands XML reader, for tests only.

``xml_io_stream<AbsorptionBand>::read`` (src/core/lbl/lbl_data.cpp:435-470) inside the ``<Map>`` of an abs_bands file and
``operator>>(std::istream&, line&)`` (lbl_data.cpp:52-58) with the Zeeman model (lbl_zeeman.cpp:311-319), the line-shape
model (lbl_lineshape_model.cpp:260-296; temperature::data lbl_temperature_model.cpp:28-43, sizes
lbl_temperature_model.h:17-33) and the local quantum numbers (quantum.cc:150-163).  Written with ``re`` and ``split``;
nothing is shared with arts_b200/csrc/xmlbands.cu.
"""
import re
from fractions import Fraction

MODEL_SIZE = {"T0": 1, "T1": 2, "T2": 3, "T3": 2, "T4": 3, "T5": 2, "AER": 4, "DPL": 4}


def read_bands(text):
    """-> list of dicts {isot, global_qn, lineshape, cutoff_type, cutoff_value, lines: [dict]}"""
    m = re.search(r'<Map\s+([^>]*)>', text)
    attrs = dict(re.findall(r'(\w+)="([^"]*)"', m.group(1)))
    assert attrs["type"] == "AbsorptionBand"
    n = int(attrs["nelem"])
    pat = re.compile(r'<QuantumIdentifier[^>]*>(.*?)</QuantumIdentifier>\s*<AbsorptionBand\s+([^>]*)>(.*?)</AbsorptionBand>', re.S)
    bands = []
    for qid, battrs, body in pat.findall(text):
        a = dict(re.findall(r'(\w+)="([^"]*)"', battrs))
        q = qid.split()
        tok = body.split()
        pos = 0

        def nxt():
            nonlocal pos
            pos += 1
            return tok[pos - 1]

        lines = []
        for _ in range(int(a["nelem"])):
            ln = {k: float(nxt()) for k in ("f0", "a", "e0", "gu", "gl")}
            ln["z_on"] = int(nxt()) != 0
            ln["z_gu"], ln["z_gl"] = float(nxt()), float(nxt())
            ln["T0"] = float(nxt())
            ln["broadeners"] = []
            for _ in range(int(nxt())):
                sp = nxt()
                models = {}
                for _ in range(int(nxt())):
                    var, typ = nxt(), nxt()
                    size = MODEL_SIZE[typ] if typ in MODEL_SIZE else int(nxt())
                    models[var] = (typ, [float(nxt()) for _ in range(size)])
                ln["broadeners"].append((sp, models))
            ln["qn"] = {}
            for _ in range(int(nxt())):
                key, up, lo = nxt(), nxt(), nxt()
                ln["qn"][key] = (up, lo)
            lines.append(ln)
        assert pos == len(tok), "nelem does not match the lines"
        bands.append(dict(isot=q[0], global_qn=q[1:], lineshape=a["lineshape"], cutoff_type=a["cutoff_type"],
                          cutoff_value=float(a["cutoff_value"]), lines=lines))
    assert len(bands) == n
    return bands


def two_j(s):
    f = 2 * Fraction(s)
    assert f.denominator == 1
    return int(f)
