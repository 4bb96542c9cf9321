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


def test_stage2_workspace_adopts_K_from_level_workspaces(wsm):
    """The C-ABI pieces of the level split (ab200_path_create_stage2, ab200_path_adopt_K), as bench.py uses them under torchrun:
    two workspaces sum the even and the odd levels for all frequencies, their K rows are copied (torch, device to device) into
    a Stokes-only workspace of a frequency block, and the radiances equal the ordinary one-workspace run bit for bit - for a
    scalar case (the scalar Stokes instantiation must be chosen from the catalog, not from who wrote K) and a Zeeman case."""
    torch = pytest.importorskip("torch")
    from arts_b200 import _abi as abi
    from arts_b200 import shard

    dev = torch.device("cuda", 0)
    for c in (synth.case_c2(lines_per_species=300, nf=777, np_=7, bands_per_species=3), synth.case_c3(nf=38 * 9, np_=5, los=(120.0, 30.0))):
        I1, _, K1 = wsm.spectral_radClearskyEmission(c.cat, c.f, c.atm, c.r, c.I_bkg, return_propmat=True)
        cat = wsm.Catalog(c.cat)
        off, cnt = 128, c.nf - 128 - 7  # a frequency block in the middle of the grid
        p2 = wsm.Path(cat, cnt, c.np_, 0, stage2_only=True)
        p2.upload(c.f[off:off + cnt], c.atm, c.r, c.I_bkg[off:off + cnt])
        K2 = shard.as_torch(p2.device_ptr(1), (c.np_, p2.k_pitch, 7), dev)
        a = c.atm
        for part in (0, 1):
            lv = list(range(part, c.np_, 2))
            pick = lambda x: None if x is None else np.ascontiguousarray(x[lv])  # noqa: E731
            sub = abi.AtmPath(T=pick(a.T), P=pick(a.P), vmr=pick(a.vmr), isorat=pick(a.isorat), Q=pick(a.Q), mag=pick(a.mag), los=pick(a.los))
            p1 = wsm.Path(cat, c.nf, len(lv), 0)
            p1.upload(c.f, sub, np.zeros(max(len(lv) - 1, 1)), None)
            p1.run_propmat()
            p1.sync()
            Kp = shard.as_torch(p1.device_ptr(1), (len(lv), p1.k_pitch, 7), dev)
            K2[lv, :cnt, :] = Kp[:, off:off + cnt, :]
            torch.cuda.synchronize()
            p1.close()
        with pytest.raises(wsm.Ab200Error, match="no line-sum buffers"):
            p2.run_propmat()
        p2.adopt_K()
        p2.run_stokes()
        I = np.empty((cnt, 4)); K = np.empty((c.np_, cnt, 7))
        p2.download(I=I, K=K)
        assert np.array_equal(K, K1[:, off:off + cnt])
        assert np.array_equal(I, I1[off:off + cnt])
        p2.close(); cat.close()
