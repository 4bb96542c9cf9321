"""The AbsorptionBands XML loader (host code, no GPU needed): bit for bit against a pure-Python
restatement of the reference's reader (tests/xml_ref.py), on the reference's own fixture tests/core/nlte/nlte_lines.xml
(tests/golden/xml_bands_fixture.json) and on a synthetic file that exercises what the fixture does not (Zeeman, rational J,
ByLine cutoff, mirrored bands, POLY, inf / scientific notation, zero G2 entries)."""
import json
import os

import numpy as np
import pytest

from arts_b200 import _abi as abi
from arts_b200 import wsm
from arts_b200._lib import Ab200Error
from tests import xml_ref

HERE = os.path.dirname(__file__)
ISO = [("H2O-161", 0, 18.0106), ("O2-66", 1, 31.9898), ("O2-68", 1, 33.9941)]
NAMES = {"Nitrogen": 2, "N2": 2, "Oxygen": 1, "O2": 1, "Water": 0, "H2O": 0, "CarbonDioxide": 3, "Hydrogen": 4, "Helium": 5,
         "Bath": abi.SPECIES_BATH}
VAR = {"G0": abi.VAR_G0, "D0": abi.VAR_D0, "DV": abi.VAR_DV, "Y": abi.VAR_Y, "G": abi.VAR_G}
TM = {"T0": abi.TM_T0, "T1": abi.TM_T1, "T2": abi.TM_T2, "T3": abi.TM_T3, "T4": abi.TM_T4, "T5": abi.TM_T5, "AER": abi.TM_AER,
      "DPL": abi.TM_DPL, "POLY": abi.TM_POLY}


def _check_against_python_reader(text, cat):
    bands = xml_ref.read_bands(text)
    assert len(bands) == len(cat.band_isot)
    iso_names = [i[0] for i in ISO]
    l = e = 0
    for b, band in enumerate(bands):
        assert iso_names[cat.band_isot[b]] == band["isot"]
        assert cat.band_offset[b] == l
        assert cat.band_lineshape[b] == {"VP_LTE": abi.LINESHAPE_VP_LTE, "VP_LTE_MIRROR": abi.LINESHAPE_VP_LTE_MIRROR}.get(
            band["lineshape"], abi.LINESHAPE_OTHER)
        assert cat.band_cutoff_type[b] == (1 if band["cutoff_type"] == "ByLine" else 0)
        assert cat.band_cutoff_value[b] == band["cutoff_value"]
        for ln in band["lines"]:
            for k in ("f0", "a", "e0", "gu", "gl", "T0", "z_gu", "z_gl"):
                assert getattr(cat, k)[l] == ln[k], (b, l, k)  # bit for bit: both sides parse the same decimal strings
            assert bool(cat.z_on[l]) == ln["z_on"]
            if "J" in ln["qn"]:
                assert cat.two_Ju[l] == xml_ref.two_j(ln["qn"]["J"][0]) and cat.two_Jl[l] == xml_ref.two_j(ln["qn"]["J"][1])
            assert cat.ls_offset[l] == e
            for sp, models in ln["broadeners"]:
                assert cat.ls_species[e] == NAMES[sp]
                want_type = np.full(abi.NVAR, abi.TM_ABSENT)
                want_X = np.zeros((abi.NVAR, 4))
                for var, (typ, x) in models.items():
                    if var in VAR:
                        want_type[VAR[var]] = TM[typ]
                        want_X[VAR[var], :len(x)] = x
                assert np.array_equal(cat.ls_type[e], want_type)
                assert np.array_equal(cat.ls_X[e], want_X)
                e += 1
            l += 1
    assert cat.band_offset[len(bands)] == l == cat.n_lines and cat.ls_offset[l] == e == len(cat.ls_species)


def test_reference_fixture():
    g = json.load(open(os.path.join(HERE, "golden", "xml_bands_fixture.json")))
    cat = wsm.abs_bandsReadXML(text=g["text"], isotopologues=ISO, species_names=NAMES, n_species=6)
    _check_against_python_reader(g["text"], cat)
    b0 = g["band0"]
    assert cat.n_lines == 3 and len(cat.band_isot) == 3 and ISO[cat.band_isot[0]][0] == b0["isot"]
    for k in ("f0", "a", "e0", "gu", "gl", "T0"):
        assert getattr(cat, k)[0] == b0[k]
    assert cat.ls_offset[1] - cat.ls_offset[0] == b0["n_broadeners"] and cat.ls_species[0] == NAMES[b0["first_broadener"]]
    assert cat.ls_type[0, abi.VAR_G0] == TM[b0["G0_type"]] and cat.ls_X[0, abi.VAR_G0, 0] == b0["G0_X0"]
    assert cat.ls_X[0, abi.VAR_G0, 1] == b0["G0_X1"]
    assert cat.ls_type[0, abi.VAR_D0] == TM[b0["D0_type"]] and cat.ls_X[0, abi.VAR_D0, 0] == b0["D0_X0"]
    assert not cat.z_on.any()


SYNTH = """<?xml version="1.0"?>
<arts format="ascii" version="1">
<Map type="AbsorptionBand" key="QuantumIdentifier" nelem="3">
<QuantumIdentifier version="1"> O2-66 ElecStateLabel X X Lambda 0 0 S 1 1 v 0 0 </QuantumIdentifier>
<AbsorptionBand lineshape="VP_LTE" cutoff_type="ByLine" cutoff_value="7.5e11" nelem="2">
56264774626.9 1.1e-10 2.0e-22 7 9 1 0.5 0.25 296 1 Bath 1 G0 AER 1 2 3 4 1 J 3 4
118750348044.712 4.479289583303983e-09 0 3 1 1 1.0011 -1.0011e+00 296 2 O2 3 G0 T1 16340.5 0.754 Y T4 1.1e-06 -2.2e-07 0.8 G2 T0 0 Bath 2 G0 T1 16000 +0.7 DV POLY 3 1e-3 2e-6 -3e-9 2 J 1 0 N 1 1
</AbsorptionBand>
<QuantumIdentifier version="1"> O2-68 </QuantumIdentifier>
<AbsorptionBand nelem="1" cutoff_value="inf" cutoff_type="None" lineshape="VP_LTE_MIRROR">
6.0e10 1e-9 1e-21 4 6 1 2.0 1.5 250.5 1 Nitrogen 2 G0 DPL 1e4 0.7 2e3 0.3 D0 T2 -100 0.5 0.01 1 J 3/2 5/2
</AbsorptionBand>
<QuantumIdentifier version="1"> H2O-161 J 1 1 </QuantumIdentifier>
<AbsorptionBand lineshape="VP_ECS_MAKAROV" cutoff_type="None" cutoff_value="0" nelem="0">
</AbsorptionBand>
</Map>
</arts>
"""


def test_synthetic_file_with_everything_the_fixture_lacks():
    cat = wsm.abs_bandsReadXML(text=SYNTH, isotopologues=ISO, species_names=NAMES, n_species=6)
    _check_against_python_reader(SYNTH, cat)
    assert list(cat.band_offset) == [0, 2, 3, 3] and cat.band_lineshape[2] == abi.LINESHAPE_OTHER
    assert cat.z_on.tolist() == [1, 1, 1] and cat.two_Ju.tolist() == [6, 2, 3] and cat.two_Jl.tolist() == [8, 0, 5]
    assert np.isinf(cat.band_cutoff_value[1]) and cat.band_cutoff_type.tolist() == [1, 0, 0]
    assert cat.ls_type[2, abi.VAR_DV] == abi.TM_POLY and cat.ls_X[2, abi.VAR_DV].tolist() == [1e-3, 2e-6, -3e-9, 0.0]
    # the same through a file
    import tempfile

    with tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False) as f:
        f.write(SYNTH)
    try:
        cat2 = wsm.abs_bandsReadXML(file=f.name, isotopologues=ISO, species_names=NAMES, n_species=6)
    finally:
        os.unlink(f.name)
    for k in ("f0", "a", "e0", "gu", "gl", "T0", "ls_X", "ls_type", "ls_species", "ls_offset", "band_offset", "two_Ju", "z_gu"):
        assert np.array_equal(getattr(cat, k), getattr(cat2, k)), k


def test_error_behaviour():
    def read(text, iso=ISO, names=NAMES):
        return wsm.abs_bandsReadXML(text=text, isotopologues=iso, species_names=names, n_species=6)

    with pytest.raises(Ab200Error) as e:
        read(SYNTH.replace("O2-68", "O2-67"))
    assert e.value.code == abi.ERR_INVALID and "unknown isotopologue" in str(e.value)
    with pytest.raises(Ab200Error) as e:
        read(SYNTH.replace("Nitrogen", "Argon"))
    assert e.value.code == abi.ERR_INVALID and "unknown broadener" in str(e.value)
    with pytest.raises(Ab200Error) as e:
        read(SYNTH.replace('nelem="2"', 'nelem="3"'))  # the reference's stream would run into the closing tag as well
    assert e.value.code == abi.ERR_INVALID
    with pytest.raises(Ab200Error) as e:
        read(SYNTH.replace("G2 T0 0", "G2 T0 12.5"))
    assert e.value.code == abi.ERR_UNSUPPORTED
    with pytest.raises(Ab200Error) as e:
        read(SYNTH.replace("DV POLY 3 1e-3 2e-6 -3e-9", "DV POLY 5 1e-3 2e-6 -3e-9 1e-12 1e-15"))
    assert e.value.code == abi.ERR_UNSUPPORTED
    with pytest.raises(Ab200Error) as e:
        read(SYNTH.replace("1 J 3 4", "1 N 3 4"))  # Zeeman on without a local J: qn.at(J) throws in the reference
    assert e.value.code == abi.ERR_INVALID
    with pytest.raises(Ab200Error) as e:
        read(SYNTH.replace("G0 AER 1 2 3 4", "G0 AER 1 2 3"))
    assert e.value.code == abi.ERR_INVALID
    with pytest.raises(Ab200Error) as e:
        read('<arts><Vector nelem="3"> 1 2 3 </Vector></arts>')
    assert e.value.code == abi.ERR_INVALID
