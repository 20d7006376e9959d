"""Multi-GPU sampling: one process per GPU, the batch split by seed, ONE collective at the end.

Mirrors what sample_and_save.py:25-46 obtains from accelerate (`split_batches=True`: each batch of
seeds is split contiguously across ranks; per-sample generators make sample i depend on seed i
only, models/diffusion/base.py:81-85).  There is no communication inside the 256-step loop; the
finished samples are all-gathered once (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int):
    """Contiguous split of n items; the first n % world_size ranks get one extra."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(seeds: Sequence[int], world_size: int, rank: int) -> List[int]:
    lo, hi = shard_bounds(len(seeds), world_size, rank)
    return list(seeds[lo:hi])


def gather_samples(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather per-rank sample tensors [n_r, ...] (n_r may differ by one) into [sum n_r, ...]."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    nmax = max(counts)
    pad = torch.zeros(nmax, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    flat = torch.empty(world * nmax, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, pad, group=group)
    out = flat.view(world, nmax, *local.shape[1:])
    if all(c == nmax for c in counts):
        return flat
    return torch.cat([out[r, : counts[r]] for r in range(world)], dim=0)


def sample_sharded(sample_fn: Callable[[List[int]], torch.Tensor], seeds: Sequence[int], group=None) -> torch.Tensor:
    """Run `sample_fn(local_seeds) -> [n_local, ...]` on this rank's contiguous shard of `seeds` and
    return the full batch (same order as `seeds`) on every rank."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    local = shard_seeds(seeds, world, rank)
    counts = [len(shard_seeds(seeds, world, r)) for r in range(world)]
    if len(local) > 0:
        out = sample_fn(local)
        assert out.shape[0] == len(local)
    else:
        out = None   # fewer seeds than ranks (e.g. a final partial batch): nothing to sample here
    if world > 1 and min(counts) == 0:
        # ranks without work still have to enter the collective with the right trailing shape / dtype:
        # take it from the lowest rank that has a sample
        src = next(r for r in range(world) if counts[r] > 0)
        meta = [None]
        if rank == src:
            meta = [(tuple(out.shape[1:]), out.dtype)]
        dist.broadcast_object_list(meta, src=src, group=group)
        if out is None:
            shape, dtype = meta[0]
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
            out = torch.empty((0,) + shape, dtype=dtype, device=dev)
    elif out is None:
        raise ValueError("sample_sharded needs at least one seed")
    return gather_samples(out, counts, group)
