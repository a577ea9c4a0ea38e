"""Batch sharding of independent PBS / gate evaluations across the GPUs of one box.

Every ciphertext's bootstrap is independent (bootstrapping.rs:58-65 reads only its own ciphertext, the
shared read-only BootstrappingKey and a LUT), so the batch is split into contiguous index ranges, one per
rank (one process per GPU), with the keys REPLICATED on every GPU.  The only collectives are the input
scatter and the result gather (and one all-gather per level of a layered circuit); there is no
collective inside the data path.  Works with any torch.distributed backend: NCCL with CUDA tensors on
the GPU box, gloo with CPU tensors in the CPU-only test-suite (the compute function is injected).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """Contiguous balanced split: the first `total % world` ranks get one extra element."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def _dist():
    import torch.distributed as dist
    return dist


def world_info(group=None):
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def scatter_rows(root_tensor, row_shape: Sequence[int], total: int, device, dtype, src: int = 0, group=None):
    """Rank `src` holds [total, *row_shape]; every rank receives its shard [n_r, *row_shape].

    Shards are padded to the largest shard so that the collective uses equal-size buffers (NCCL scatter
    requirement); the padding rows are dropped on receipt.
    """
    import torch
    dist = _dist()
    rank, world = world_info(group)
    lo, hi = shard_range(total, rank, world)
    if world == 1:
        return root_tensor[lo:hi]
    pad = max(shard_sizes(total, world))
    recv = torch.empty((pad, *row_shape), dtype=dtype, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for r in range(world):
            a, b = shard_range(total, r, world)
            c = torch.zeros((pad, *row_shape), dtype=dtype, device=device)
            c[: b - a] = root_tensor[a:b]
            chunks.append(c)
    dist.scatter(recv, chunks, src=src, group=group)
    return recv[: hi - lo]


def gather_rows(local, total: int, dst: int = 0, group=None):
    """Inverse of scatter_rows: rank `dst` returns [total, ...], the others None."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    pad = max(shard_sizes(total, world))
    send = torch.zeros((pad, *local.shape[1:]), dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: n] for r, n in enumerate(shard_sizes(total, world))], dim=0)


def all_gather_rows(local, total: int, group=None):
    """Every rank gets the concatenation [total, ...] of all shards (one per circuit level)."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    pad = max(shard_sizes(total, world))
    send = torch.zeros((pad, *local.shape[1:]), dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    bufs = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(bufs, send, group=group)
    return torch.cat([bufs[r][: n] for r, n in enumerate(shard_sizes(total, world))], dim=0)


def bootstrap_sharded(compute: Callable, lwe_in_root, total: int, row_len: int, device, dtype, src: int = 0, group=None):
    """scatter -> compute(local_shard) -> gather.  `compute` maps [b, row_len] -> [b, row_len]."""
    local = scatter_rows(lwe_in_root, (row_len,), total, device, dtype, src, group)
    out = compute(local) if local.shape[0] else local
    return gather_rows(out, total, src, group)
