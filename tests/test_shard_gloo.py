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
