"""A catalog read from the reference's abs_bands XML fixture (tests/golden/xml_bands_fixture.json) and from the synthetic
file of tests/test_xml_bands.py runs through the GPU path against the oracle on the same arrays."""
import json
import os

import numpy as np
import pytest

from arts_b200._abi import AtmPath
from tests.conftest import assert_propmat_close
from tests.test_xml_bands import ISO, NAMES, SYNTH

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(__file__)


def _atm(np_, n_species, n_isot, mag=False):
    z = np.linspace(0.0, 40.0, np_)
    T = 288.0 - 1.5 * z
    P = 1.0e5 * np.exp(-z / 7.5)
    vmr = np.tile(np.array([0.01, 0.21, 0.78, 4e-4, 5e-7, 5e-6])[:n_species], (np_, 1))
    return AtmPath(T=T, P=P, vmr=vmr, isorat=np.full((np_, n_isot), 0.99), Q=np.tile(170.0 * (T / 296.0)[:, None] ** 1.5, (1, n_isot)),
                   dQdT=np.tile(1.5 * 170.0 / 296.0 * (T / 296.0)[:, None] ** 0.5, (1, n_isot)),
                   mag=np.tile(np.array([1e-5, 3e-5, -2e-5]), (np_, 1)) if mag else None,
                   los=np.tile(np.array([140.0, 30.0]), (np_, 1)))


def test_reference_fixture_catalog_on_the_gpu(wsm, orc):
    g = json.load(open(os.path.join(HERE, "golden", "xml_bands_fixture.json")))
    cat = wsm.abs_bandsReadXML(text=g["text"], isotopologues=ISO, species_names=NAMES, n_species=6)
    atm = _atm(5, 6, len(ISO))
    f = np.linspace(400e9, 1300e9, 1500)
    tg = (("T",), ("VMR", 0))
    Kr, dKr = orc.propmat_levels(cat, f, atm, targets=tg)
    K, dK = wsm.spectral_propmat_pathFromPath(cat, f, atm, jac_targets=tg)
    assert Kr[..., 0].max() > 0
    assert_propmat_close(K, Kr)
    for q in range(2):
        sc = np.abs(dKr[:, q]).reshape(-1, 7).max(axis=0)
        assert (np.abs(dK[:, q] - dKr[:, q]).reshape(-1, 7).max(axis=0) <= 2e-7 * sc + 1e-300).all()


def test_synthetic_xml_catalog_with_zeeman_on_the_gpu(wsm, orc):
    text = SYNTH.replace('lineshape="VP_ECS_MAKAROV"', 'lineshape="VP_LTE"').replace(
        'lineshape="VP_LTE_MIRROR"', 'lineshape="VP_LTE"')  # a mirrored band cannot carry a cutoff neighbour-free test here
    cat = wsm.abs_bandsReadXML(text=text, isotopologues=ISO, species_names=NAMES, n_species=6)
    atm = _atm(4, 6, len(ISO), mag=True)
    f = np.concatenate([118750348044.712 + np.linspace(-3e6, 3e6, 301), np.linspace(50e9, 70e9, 200)])
    f.sort()
    Kr, _ = orc.propmat_levels(cat, f, atm)
    K, _ = wsm.spectral_propmat_pathFromPath(cat, f, atm)
    assert np.abs(Kr[..., 1:]).max() > 0, "polarised components present"
    assert_propmat_close(K, Kr, atol_scale=1e-11)
