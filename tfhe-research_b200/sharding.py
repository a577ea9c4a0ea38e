"""Batch sharding of independent PBS / gate evaluations across the GPUs of one box (one process per GPU).

Every ciphertext's bootstrap is independent (bootstrapping.rs:58-65 reads only its own ciphertext, the
shared read-only BootstrappingKey and a LUT), so the batch is split into contiguous index ranges, one per
rank, with the keys REPLICATED on every GPU.  The only communication is the input scatter and the result
gather (and one all-gather per level of a layered circuit); there is no collective inside the data path.
Works with any torch.distributed backend: NCCL with CUDA tensors on the GPU box, gloo with CPU tensors in
the CPU-only test-suite (the compute function is injected).

Scatter and gather are batched point-to-point transfers of exactly the shard's rows (`batch_isend_irecv`:
one NCCL group): the root sends row slices of its tensor and receives straight into row slices of the
result -- no padding, no staging copies, no concatenation.  (The one-process form of the same thing, for a
host that is not Python, is `tfhe_mgpu_*` in include/tfhe_b200.h.)
"""
from __future__ import annotations

from typing import Callable, List, Sequence


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


def _run(ops):
    if ops:
        for req in _dist().batch_isend_irecv(ops):
            req.wait()


def scatter_rows(root_tensor, row_shape: Sequence[int], total: int, device, dtype, src: int = 0, group=None):
    """Rank `src` holds [total, *row_shape]; every rank receives its shard [n_r, *row_shape] (the root: a view)."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    lo, hi = shard_range(total, rank, world)
    if world == 1:
        return root_tensor[lo:hi]
    if rank == src:
        ops = []
        for r in range(world):
            a, b = shard_range(total, r, world)
            if r != src and b > a:
                ops.append(dist.P2POp(dist.isend, root_tensor[a:b], r, group))
        _run(ops)
        return root_tensor[lo:hi]
    recv = torch.empty((hi - lo, *row_shape), dtype=dtype, device=device)
    if hi > lo:
        _run([dist.P2POp(dist.irecv, recv, src, group)])
    return recv


def gather_rows(local, total: int, dst: int = 0, group=None, out=None):
    """Inverse of scatter_rows: rank `dst` returns [total, ...] (`out` if given), the others None."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    if rank != dst:
        if local.shape[0]:
            _run([dist.P2POp(dist.isend, local.contiguous(), dst, group)])
        return None
    if out is None:
        out = torch.empty((total, *local.shape[1:]), dtype=local.dtype, device=local.device)
    ops = []
    for r in range(world):
        a, b = shard_range(total, r, world)
        if r == dst:
            out[a:b] = local
        elif b > a:
            ops.append(dist.P2POp(dist.irecv, out[a:b], r, group))
    _run(ops)
    return out


def all_gather_rows(local, total: int, group=None):
    """Every rank gets the concatenation [total, ...] of all shards (one per circuit level)."""
    import torch
    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    out = torch.empty((total, *local.shape[1:]), dtype=local.dtype, device=local.device)
    if total % world == 0:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)   # equal shards: one collective, straight into `out`
        return out
    sizes = shard_sizes(total, world)
    pad = max(sizes)
    send = torch.zeros((pad, *local.shape[1:]), dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    buf = torch.empty((world * pad, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, send, group=group)
    for r, n in enumerate(sizes):
        a, _ = shard_range(total, r, world)
        out[a:a + n] = buf[r * pad:r * pad + n]
    return out


def bootstrap_sharded(compute: Callable, lwe_in_root, total: int, row_len: int, device, dtype, src: int = 0, group=None, out=None):
    """scatter -> compute(local_shard) -> gather.  `compute` maps [b, row_len] -> [b, row_len]."""
    local = scatter_rows(lwe_in_root, (row_len,), total, device, dtype, src, group)
    res = compute(local) if local.shape[0] else local
    return gather_rows(res, total, src, group, out)
