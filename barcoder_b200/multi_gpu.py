"""Library sharding across the GPUs of one box.

One process per GPU (torchrun); every rank holds the whole packed genome and the seed index
of ITS contiguous slice of the library, searches it with the single-GPU path, and the 16-byte
hit records are gathered to rank 0 with one all_gather of counts plus grouped NCCL
send/recv of the raw records (a gather-v).  There is no collective inside the search itself:
(spacer, position) pairs are independent.  torch.distributed is plumbing only.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n, world_size, rank):
    """Contiguous, balanced slice [lo, hi) of n library rows for `rank`."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DeviceBuffer:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n_records):
        self.__cuda_array_interface__ = {
            "shape": (n_records, 4), "typestr": "<i4", "data": (ptr, False), "version": 3, "strides": None,
        }


def hits_as_tensor(searcher, device):
    """The searcher's device-resident hit records as an int32 [n, 4] tensor (no copy)."""
    import torch
    ptr, n = searcher.hits_device()
    if n == 0:
        return torch.empty((0, 4), dtype=torch.int32, device=device)
    return torch.as_tensor(_DeviceBuffer(ptr, n), device=device)


def gather_hits(local, group=None, dst=0):
    """Gather variable-length int32 [n_i, 4] record tensors to `dst`.
    Returns the concatenated tensor on dst (None elsewhere) and the per-rank counts."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    mine = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    counts_host = counts.cpu().tolist()
    if rank == dst:
        out = torch.empty((sum(counts_host), 4), dtype=torch.int32, device=local.device)
        ops, off = [], 0
        for r, c in enumerate(counts_host):
            if r == dst:
                out[off:off + c].copy_(local)
            elif c:
                ops.append(dist.P2POp(dist.irecv, out[off:off + c], r, group))
            off += c
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out, counts_host
    if local.shape[0]:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
            req.wait()
    return None, counts_host


def records_from_tensor(t):
    """int32 [n, 4] tensor -> structured numpy hit records on the host."""
    from ._native import HIT_DTYPE
    arr = t.cpu().numpy().astype(np.int32, copy=False)
    return np.ascontiguousarray(arr).view(np.uint32).reshape(-1, 4).copy().view(HIT_DTYPE).reshape(-1)
