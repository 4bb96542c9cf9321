"""Frequency sharding of the clear-sky path across the GPUs of one box.

The path is independent per frequency end to end (line sum per f, RTE chain per f), so the
ascending ``freq_grid`` is cut into contiguous blocks — the same split the reference uses for
its OpenMP frequency chunks, ``matpack::omp_offset_count``
(src/core/matpack/matpack_mdspan_algorithm.cc:4-18, called from src/m_lbl.cc:273) — one block
per rank, catalog and atmosphere replicated.  There is no data-path collective in the compute;
the only exchange is the gather of ``spectral_rad`` [nf, 4] (and nothing else) at the end.

``torch.distributed`` is the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU
tests); the arrays being gathered are the library's own device buffers, wrapped zero-copy.
"""
from __future__ import annotations

import numpy as np


def frequency_ranges(nf: int, n: int) -> list[tuple[int, int]]:
    """``[(offset, nelem)] * n`` exactly as matpack::omp_offset_count(nf, n): ``nf // n`` elements
    per block, the remainder to the last block."""
    if n < 1:
        raise ValueError("need at least one shard")
    if n == 1:
        return [(0, nf)]
    dn = nf // n
    out = [(i * dn, dn) for i in range(n - 1)]
    out.append(((n - 1) * dn, nf - (n - 1) * dn))
    return out


def max_block(nf: int, n: int) -> int:
    return max(c for _, c in frequency_ranges(nf, n))


class DeviceArray:
    """Zero-copy view of a library-owned device buffer for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
            "strides": None,
        }


def as_torch(ptr: int, shape, device):
    import torch

    return torch.as_tensor(DeviceArray(ptr, shape), device=device)


def gather_spectral_rad(local, nf: int, group=None, dst: int | None = None):
    """Gathers the per-rank blocks ``local`` [nelem_rank, 4] (torch tensor on the rank's device, or
    CPU tensor with gloo) into the full ``spectral_rad`` [nf, 4].

    Blocks are padded to the largest block so that one ``all_gather_into_tensor`` (NCCL
    AllGather over NVLink/NVSwitch) moves everything; with ``dst`` set only that rank assembles
    and returns the result (others return None).  The value at frequency j never depends on
    the number of ranks: blocks are copied, not reduced.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ranges = frequency_ranges(nf, world)
    off, cnt = ranges[rank]
    if tuple(local.shape) != (cnt, 4):
        raise ValueError(f"rank {rank}: local block has shape {tuple(local.shape)}, expected {(cnt, 4)}")
    blk = max(c for _, c in ranges)
    send = local
    if cnt != blk:
        send = torch.zeros((blk, 4), dtype=local.dtype, device=local.device)
        send[:cnt] = local
    recv = torch.empty((world * blk, 4), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if dst is not None and rank != dst:
        return None
    if all(c == blk for _, c in ranges):
        return recv[:nf]
    out = torch.empty((nf, 4), dtype=local.dtype, device=local.device)
    for r, (o, c) in enumerate(ranges):
        out[o:o + c] = recv[r * blk:r * blk + c]
    return out


def shard_case(case, rank: int, world: int):
    """The frequency shard of a synthetic ``Case`` for one rank (catalog / atmosphere shared)."""
    import copy

    off, cnt = frequency_ranges(case.nf, world)[rank]
    c = copy.copy(case)
    c.f = np.ascontiguousarray(case.f[off:off + cnt])
    c.I_bkg = np.ascontiguousarray(case.I_bkg[off:off + cnt])
    return c, off, cnt


# ---------------------------------------------------------------------------
# levels for the line sum, frequencies for the Stokes chain: one transpose of K between them
# ---------------------------------------------------------------------------
def level_sets(np_: int, n: int) -> list[list[int]]:
    """Levels dealt round-robin: rank r sums levels r, r + n, r + 2n, ... (the split of ab200_multi_*, multi.cu)."""
    return [list(range(r, np_, n)) for r in range(n)]


class LevelExchange:
    """The one exchange of the level split: rank r holds ``K1`` [levels of r][k_pitch_all][7] for ALL frequencies and needs
    ``K2`` [np][k_pitch_local][7] for ITS frequency block — an all-to-all of K (NCCL over NVLink / NVSwitch on the GPU box,
    gloo in the CPU tests).  Rows are copied, never reduced, and K at a (frequency, level) does not depend on the partition,
    so the radiances are bit-identical to the frequency split and to one GPU.

    Why: the line records and the cluster moments of the far-field sums are work per (line, level); a frequency shard
    repeats them for every rank (25 of 60 ms per configs[3] shard), a level shard does not.
    """

    def __init__(self, np_: int, nf_total: int, rank: int, world: int, group=None):
        import torch

        self.np_, self.nf, self.rank, self.world, self.group = int(np_), int(nf_total), rank, world, group
        self.ranges = frequency_ranges(self.nf, world)
        if len({c for _, c in self.ranges}) != 1:
            raise ValueError("the level exchange needs equal frequency blocks (nf divisible by the number of ranks)")
        self.cnt = self.ranges[0][1]
        self.levels = level_sets(self.np_, world)
        self.mine = self.levels[rank]
        self.out_split = [len(l) for l in self.levels]            # rows received from every rank
        self.in_split = [len(self.mine)] * world                  # rows sent to every rank
        self.level_of_row = torch.tensor([l for ls in self.levels for l in ls], dtype=torch.long)

    def exchange(self, K1, K2):
        """K1 [len(mine), >= nf, 7] -> K2 [np, >= cnt, 7] (views of the library's device buffers; extra columns are padding)."""
        import torch
        import torch.distributed as dist

        n_mine, cnt, w = len(self.mine), self.cnt, self.world
        if self.level_of_row.device != K1.device:
            self.level_of_row = self.level_of_row.to(K1.device)
        # [levels][rank blocks][cnt][7] -> [rank blocks][levels][cnt][7]: the rows for rank e are contiguous
        send = K1[:, : w * cnt, :].reshape(n_mine, w, cnt, 7).permute(1, 0, 2, 3).contiguous().view(w * n_mine, cnt * 7)
        recv = torch.empty((self.np_, cnt * 7), dtype=K1.dtype, device=K1.device)
        if n_mine == 0 and all(s == 0 for s in self.out_split):
            return
        dist.all_to_all_single(recv, send, output_split_sizes=self.out_split, input_split_sizes=self.in_split, group=self.group)
        K2[:, :cnt, :].index_copy_(0, self.level_of_row, recv.view(self.np_, cnt, 7))


# ---------------------------------------------------------------------------
# second axis: a batch of propagation paths (BASELINE configs[4]) sharded over the ranks
# ---------------------------------------------------------------------------
def path_ranges(n_paths: int, n: int) -> list[tuple[int, int]]:
    """Contiguous blocks of the simulations of ``measurement_vecFromSensor`` (src/m_rad.cc:321-362), one per rank,
    the same split as the frequency axis."""
    return frequency_ranges(n_paths, n)


def reduce_measurement(y, J, n_paths: int | None = None, group=None, ordered: bool = False):
    """Combines the per-rank channel sums ``y`` [M] and ``J`` [M, nx] of a path-sharded batch.

    The reference adds one simulation after the other into ``measurement_vec`` (:346-351).  ``ordered=False``: one
    all-reduce (NCCL over NVLink / NVSwitch) of the rank sums - the cheapest exchange, the result depends on the number
    of ranks in the last bits.  ``ordered=True``: ``y`` [n_local, M] and ``J`` [n_local, M, nx] hold the contribution of
    every path of the rank; blocks are gathered and summed in path order on every rank, so the value is the same for any
    number of ranks (at the price of moving n_paths * M * (nx + 1) doubles).
    """
    import torch
    import torch.distributed as dist

    if not ordered:
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(J, op=dist.ReduceOp.SUM, group=group)
        return y, J
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if n_paths is None:
        raise ValueError("ordered reduction needs the total number of paths")
    ranges = path_ranges(n_paths, world)
    off, cnt = ranges[rank]
    if y.shape[0] != cnt or J.shape[0] != cnt:
        raise ValueError(f"rank {rank}: {y.shape[0]} path contributions, expected {cnt}")
    blk = max(c for _, c in ranges)
    M, nx = J.shape[1], J.shape[2]
    send = torch.zeros((blk, M, nx + 1), dtype=J.dtype, device=J.device)
    send[:cnt, :, :nx] = J
    send[:cnt, :, nx] = y
    recv = torch.empty((world * blk, M, nx + 1), dtype=J.dtype, device=J.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    tot = torch.zeros((M, nx + 1), dtype=J.dtype, device=J.device)
    for r, (_, c) in enumerate(ranges):  # path order: rank blocks are contiguous and ascending
        for i in range(c):
            tot += recv[r * blk + i]
    return tot[:, nx].clone(), tot[:, :nx].clone()
