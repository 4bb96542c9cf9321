"""One host process, several GPUs behind the C ABI (ab200_multi_*): the call a shim inside the reference's single process
makes.  The grid is dealt over the devices in 512-frequency blocks; results must be bit-identical to the one-device call
(forward spectra, propagation matrix incl. `+=`, Jacobian rows of real lines), with and without ByLine cutoffs (the line
selection must see the bounds of the whole grid).  On a one-GPU box the same device is used twice — the code path (two
catalog replicas, two worker threads, block dealing, scatter into the caller's arrays) is the same."""
import numpy as np
import pytest

from arts_b200 import synth

pytestmark = pytest.mark.gpu


def _devices(wsm, want):
    n = wsm.device_count()
    return list(range(want)) if n >= want else [i % n for i in range(want)]


@pytest.mark.parametrize("ndev", [2, 3])
@pytest.mark.parametrize("cutoff", [None, 3e9])
def test_multi_device_clearsky_is_bitwise_the_single_device_result(wsm, ndev, cutoff):
    c = synth.case_c2(lines_per_species=300, nf=5 * 512 + 137, np_=9, bands_per_species=3)
    if cutoff:
        c.cat.band_cutoff_type[:] = 1
        c.cat.band_cutoff_value[:] = cutoff
    tg = (("T",), ("VMR", 0))
    I1, dI1, K1 = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1, return_propmat=True)
    m = wsm.MultiDevice(c.cat, devices=_devices(wsm, ndev))
    assert m.n_devices == ndev
    for _ in range(2):  # cached workspaces on the second call
        I, dI, K = wsm.spectral_radClearskyEmission(m, c.f, c.atm, c.r, c.I_bkg, jac_targets=tg, hse_derivative=1, return_propmat=True)
        assert np.array_equal(K, K1)
        assert np.array_equal(I, I1)
        assert np.array_equal(dI, dI1)
    m.close()


def test_multi_device_propmat_accumulates_and_takes_level_grids(wsm):
    c = synth.tiny_case(nl=200, nf=1400, np_=5)
    fp = np.ascontiguousarray(np.stack([c.f * (1.0 + 1e-6 * i) for i in range(c.np_)]))  # one grid per level
    tg = (("VMR", 0),)
    K1 = np.full((c.np_, c.nf, 7), 0.5)
    dK1 = np.full((c.np_, 1, c.nf, 7), -2.0)
    wsm.spectral_propmat_pathFromPath(c.cat, fp, c.atm, jac_targets=tg, out=K1, out_jac=dK1, accumulate=True)
    m = wsm.MultiDevice(c.cat, devices=_devices(wsm, 2))
    K = np.full((c.np_, c.nf, 7), 0.5)
    dK = np.full((c.np_, 1, c.nf, 7), -2.0)
    wsm.spectral_propmat_pathFromPath(m, fp, c.atm, jac_targets=tg, out=K, out_jac=dK, accumulate=True)
    assert np.array_equal(K, K1) and np.array_equal(dK, dK1)
    assert np.all(K[..., 1:] == 0.5)
    m.close()


def test_multi_device_zeeman_and_errors(wsm):
    c = synth.case_c3(nf=38 * 40, np_=4, los=(120.0, 30.0))
    I1, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    m = wsm.MultiDevice(c.cat, devices=_devices(wsm, 2))
    I, _ = wsm.spectral_radClearskyEmission(m, c.f, c.atm, c.r, c.I_bkg)
    assert np.array_equal(I, I1)
    with pytest.raises(wsm.Ab200Error) as e:  # the workers' error comes back to the caller, with the device it came from
        wsm.spectral_radClearskyEmission(m, c.f, c.atm, c.r, c.I_bkg, select_species=99)
    assert "device" in str(e.value) and "select_species" in str(e.value)
    m.close()
    with pytest.raises(wsm.Ab200Error):
        wsm.MultiDevice(c.cat, n_devices=wsm.device_count() + 1)


@pytest.mark.parametrize("ndev", [2, 3])
@pytest.mark.parametrize("cutoff", [None, 3e9])
def test_level_split_forward_is_bitwise_the_single_device_result(wsm, ndev, cutoff, monkeypatch):
    """Forward calls on a shared grid split the LEVELS of the line sum over the devices and the frequencies of the Stokes
    chain, with one transpose of K between them (device-to-device copies); 10 levels over 3 devices and a grid that is not
    a multiple of 128 exercise the ragged ends.  Same bits as one device and as the frequency split."""
    c = synth.case_c2(lines_per_species=300, nf=5 * 512 + 137, np_=10, bands_per_species=3)
    if cutoff:
        c.cat.band_cutoff_type[:] = 1
        c.cat.band_cutoff_value[:] = cutoff
    I1, _, K1 = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, return_propmat=True)
    m = wsm.MultiDevice(c.cat, devices=_devices(wsm, ndev))
    for _ in range(2):
        I, _, K = wsm.spectral_radClearskyEmission(m, c.f, c.atm, c.r, c.I_bkg, return_propmat=True)
        assert np.array_equal(K, K1)
        assert np.array_equal(I, I1)
    I, _ = wsm.spectral_radClearskyEmission(m, c.f, c.atm, c.r, c.I_bkg, rte_option="constant")
    Ic, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, rte_option="constant")
    assert np.array_equal(I, Ic)
    monkeypatch.setenv("AB200_MULTI_SPLIT", "freq")
    I, _ = wsm.spectral_radClearskyEmission(m, c.f, c.atm, c.r, c.I_bkg)
    assert np.array_equal(I, I1)
    m.close()


def test_level_split_with_far_field_sums_and_wind(wsm, monkeypatch):
    monkeypatch.setenv("AB200_FARFIELD", "2")  # far-field sums for every real segment, whatever its size
    c = synth.case_c2(lines_per_species=400, nf=3000, np_=7, bands_per_species=2)
    c.atm.wind = np.tile([10.0, -5.0, 0.5], (c.np_, 1)) * np.linspace(0.5, 1.5, c.np_)[:, None]
    c.atm.los = np.tile([140.0, 20.0], (c.np_, 1))
    I1, _ = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg)
    m = wsm.MultiDevice(c.cat, devices=_devices(wsm, 2))
    I, _ = wsm.spectral_radClearskyEmission(m, c.f, c.atm, c.r, c.I_bkg)
    assert np.array_equal(I, I1)
    m.close()
