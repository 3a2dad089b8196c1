"""Batch-sharded multi-GPU sampling (one process per GPU, torch.distributed).

Point clouds are independent (no cross-sample operation anywhere on the path), so the
batch is partitioned across ranks with NO per-step collective; the only exchange is one
all-gather of the finished clouds (NCCL over NVLink on the GPU box, gloo in the CPU tests).
The reference has no multi-GPU inference path -- this is an addition.
"""
from typing import Any, Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ``batch`` items for ``rank``; the remainder goes to the
    low ranks (sizes differ by at most one)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_kwargs(model_kwargs: Dict[str, Any], batch: int, world: int, rank: int) -> Dict[str, Any]:
    lo, hi = shard_bounds(batch, world, rank)
    out = {}
    for k, v in model_kwargs.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == batch:
            out[k] = v[lo:hi]
        elif isinstance(v, (list, tuple)) and len(v) == batch:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def _world_rank(group=None) -> Tuple[int, int]:
    """(world size, rank); a process without an initialised process group is a world of one."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_clouds(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gather the per-rank results [b_r, C, N] into [batch, C, N] on every rank
    (ragged shards are padded to the largest shard for the collective)."""
    world, _ = _world_rank(group)
    if world == 1:
        return local
    biggest = -(-batch // world)
    C, N = local.shape[1], local.shape[2]
    out = local.new_empty((world * biggest, C, N))
    if batch % world == 0:  # even shards: the local result is the send buffer
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = local.new_zeros((biggest, C, N))
    padded[: local.shape[0]] = local
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(batch, world, r)
        parts.append(out[r * biggest: r * biggest + (hi - lo)])
    return torch.cat(parts, dim=0)


def sample_sharded(sample_fn: Callable[[int, Dict[str, Any]], torch.Tensor], batch: int,
                   model_kwargs: Dict[str, Any], group=None, gather: bool = True, kwargs_are_local: bool = False,
                   local_batch: Optional[int] = None) -> torch.Tensor:
    """Run ``sample_fn(local_batch, local_kwargs)`` (e.g. ``PointCloudSampler.sample_batch``) on this
    rank's shard of a ``batch``-cloud job and return the gathered [batch, C, N] result.

    ``model_kwargs`` holds the conditioning of the WHOLE batch (sliced here with ``shard_kwargs``) unless
    ``kwargs_are_local`` says the caller already holds only its own shard (e.g. it copied just that shard to
    the device); ``local_batch`` then overrides the shard size (default: ``shard_bounds``)."""
    world, rank = _world_rank(group)
    lo, hi = shard_bounds(batch, world, rank)
    n_local = hi - lo if local_batch is None else local_batch
    kw = model_kwargs if kwargs_are_local else shard_kwargs(model_kwargs, batch, world, rank)
    local = sample_fn(n_local, kw)
    return gather_clouds(local, batch, group) if gather else local
