"""The N > 1 host logic on CPU: world_size-2 (and 3) gloo runs of the frequency sharding + gather.

Each rank computes its contiguous frequency shard of a small case with the CPU oracle (standing in
for the GPU kernels, which need a device), the blocks are gathered with the same
``shard.gather_spectral_rad`` the NCCL path uses, and the assembled spectrum must be bit-identical
to the unsharded run — the reference's thread-count invariance carried over to ranks.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arts_b200 import shard, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nf, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import oracle_lib as orc

        case = synth.tiny_case(nl=48, nf=nf, np_=5, cutoff=2e9)
        mine, off, cnt = shard.shard_case(case, rank, world)
        assert (off, cnt) == shard.frequency_ranges(nf, world)[rank]
        if cnt:
            I, _ = orc.clearsky_emission(mine.cat, mine.f, mine.atm, mine.r, mine.I_bkg)
        else:
            I = np.zeros((0, 4))
        full = shard.gather_spectral_rad(torch.from_numpy(np.ascontiguousarray(I)), nf)
        only0 = shard.gather_spectral_rad(torch.from_numpy(np.ascontiguousarray(I)), nf, dst=0)
        assert (only0 is None) == (rank != 0)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        np.save(os.path.join(tmp, f"full_{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nf", [(2, 301), (3, 64), (2, 2)])
def test_sharded_gather_is_bit_identical(tmp_path, orc, world, nf):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nf, str(tmp_path)), nprocs=world, join=True)
    case = synth.tiny_case(nl=48, nf=nf, np_=5, cutoff=2e9)
    ref, _ = orc.clearsky_emission(case.cat, case.f, case.atm, case.r, case.I_bkg)
    for r in range(world):
        got = np.load(tmp_path / f"full_{r}.npy")
        assert got.shape == (nf, 4)
        # no cutoff-window line crosses a shard edge in this fixture, so the shard-local line selection
        # (lbl_data.cpp:61-68 on the shard's own bounds) equals the global one: bitwise equality
        assert np.array_equal(got, ref), f"rank {r}"


def test_gather_rejects_wrong_block_shape():
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        with pytest.raises(ValueError):
            shard.gather_spectral_rad(torch.zeros(3, 4, dtype=torch.float64), 5)
        out = shard.gather_spectral_rad(torch.arange(20, dtype=torch.float64).reshape(5, 4), 5)
        assert torch.equal(out, torch.arange(20, dtype=torch.float64).reshape(5, 4))
    finally:
        dist.destroy_process_group()


def _path_worker(rank, world, port, n_paths, tmp):
    """Path-axis sharding (BASELINE configs[4]): every rank runs the observer epilogue of its block of paths with the
    CPU oracle standing in for the GPU, then the channel sums are combined both ways."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, cnt = shard.path_ranges(n_paths, world)[rank]
        contrib = _path_contributions(n_paths)[off:off + cnt]
        M, nx = contrib.shape[1], contrib.shape[2] - 1
        y = torch.from_numpy(np.ascontiguousarray(contrib[:, :, nx]))
        J = torch.from_numpy(np.ascontiguousarray(contrib[:, :, :nx]))
        yo, Jo = shard.reduce_measurement(y.clone(), J.clone(), n_paths=n_paths, ordered=True)
        ys, Js = shard.reduce_measurement(y.sum(0), J.sum(0))
        np.save(os.path.join(tmp, f"yo_{rank}.npy"), yo.numpy())
        np.save(os.path.join(tmp, f"Jo_{rank}.npy"), Jo.numpy())
        np.save(os.path.join(tmp, f"ys_{rank}.npy"), ys.numpy())
        np.save(os.path.join(tmp, f"Js_{rank}.npy"), Js.numpy())
    finally:
        dist.destroy_process_group()


def _path_contributions(n_paths, M=4, nx=6):
    """Deterministic stand-in for the per-path (y, Jy) of ab200_path_download_observer, badly scaled on purpose so
    that the order of the additions shows in the last bits."""
    rng = np.random.default_rng(77)
    return rng.normal(size=(n_paths, M, nx + 1)) * 10.0 ** rng.uniform(-6, 6, size=(n_paths, 1, 1))


@pytest.mark.parametrize("world,n_paths", [(2, 7), (3, 10), (2, 1)])
def test_path_sharded_measurement_reduction(tmp_path, world, n_paths):
    port = _free_port()
    mp.spawn(_path_worker, args=(world, port, n_paths, str(tmp_path)), nprocs=world, join=True)
    contrib = _path_contributions(n_paths)
    seq = np.zeros(contrib.shape[1:])
    for i in range(n_paths):  # measurement_vec[iv] += ..., one simulation after the other (src/m_rad.cc:346-351)
        seq += contrib[i]
    for r in range(world):
        yo, Jo = np.load(tmp_path / f"yo_{r}.npy"), np.load(tmp_path / f"Jo_{r}.npy")
        assert np.array_equal(yo, seq[:, -1]) and np.array_equal(Jo, seq[:, :-1]), "ordered reduction: same bits as one rank"
        ys, Js = np.load(tmp_path / f"ys_{r}.npy"), np.load(tmp_path / f"Js_{r}.npy")
        np.testing.assert_allclose(ys, seq[:, -1], rtol=1e-9, atol=1e-9 * np.abs(contrib).max())
        np.testing.assert_allclose(Js, seq[:, :-1], rtol=1e-9, atol=1e-9 * np.abs(contrib).max())
    assert shard.path_ranges(10, 3) == [(0, 3), (3, 3), (6, 4)]


def _level_worker(rank, world, port, nf, np_, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arts_b200 import _abi as abi
        from tests import oracle_lib as orc

        case = synth.tiny_case(nl=48, nf=nf, np_=np_)
        ex = shard.LevelExchange(np_, nf, rank, world)
        assert ex.mine == list(range(rank, np_, world))
        a = case.atm
        sub = abi.AtmPath(T=a.T[ex.mine], P=a.P[ex.mine], vmr=a.vmr[ex.mine], isorat=a.isorat[ex.mine], Q=a.Q[ex.mine])
        K1, _ = orc.propmat_levels(case.cat, case.f, sub)  # this rank's levels, ALL frequencies (stands in for the GPU line sum)
        pitch1, pitch2 = nf + 5, ex.cnt + 3                # the library pads its level rows; the exchange must not care
        K1p = torch.full((len(ex.mine), pitch1, 7), float("nan"), dtype=torch.float64)
        K1p[:, :nf] = torch.from_numpy(K1)
        K2p = torch.full((np_, pitch2, 7), float("nan"), dtype=torch.float64)
        ex.exchange(K1p, K2p)
        off, cnt = shard.frequency_ranges(nf, world)[rank]
        K2 = np.ascontiguousarray(K2p[:, :cnt].numpy())
        assert torch.isnan(K2p[:, cnt:]).all()
        T, L, P, dT, dL = orc.tramat(K2, None, case.r, None, "linsrc")
        J, dJ = orc.srcvec(K2, case.f[off:off + cnt], a.T, -1, 0)
        I, _ = orc.rte_emission("linsrc", T, L, P, dT, dL, J, dJ, np.ascontiguousarray(case.I_bkg[off:off + cnt]))
        full = shard.gather_spectral_rad(torch.from_numpy(np.ascontiguousarray(I)), nf)
        np.save(os.path.join(tmp, f"lvl_{rank}.npy"), full.numpy())
        np.save(os.path.join(tmp, f"K2_{rank}.npy"), K2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nf,np_", [(2, 300, 7), (3, 66, 5), (2, 64, 2)])
def test_level_split_exchange_is_bit_identical(tmp_path, orc, world, nf, np_):
    """Levels dealt over the ranks for the line sum, frequency blocks for the Stokes chain, one all-to-all of K between them
    (shard.LevelExchange, the torchrun twin of ab200_multi_*'s level split): K and the radiances equal the unsharded run bit
    for bit, with ragged level counts (7 over 2, 5 over 3) and padded level rows."""
    port = _free_port()
    mp.spawn(_level_worker, args=(world, port, nf, np_, str(tmp_path)), nprocs=world, join=True)
    case = synth.tiny_case(nl=48, nf=nf, np_=np_)
    Kref, _ = orc.propmat_levels(case.cat, case.f, case.atm)
    ref, _ = orc.clearsky_emission(case.cat, case.f, case.atm, case.r, case.I_bkg)
    for r in range(world):
        off, cnt = shard.frequency_ranges(nf, world)[r]
        assert np.array_equal(np.load(tmp_path / f"K2_{r}.npy"), Kref[:, off:off + cnt]), f"K of rank {r}"
        assert np.array_equal(np.load(tmp_path / f"lvl_{r}.npy"), ref), f"rank {r}"
