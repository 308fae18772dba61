"""Data-parallel plumbing (new relative to the single-GPU reference).

The path shards by samples: every rank owns a full replica and B images; the ONE exchange step is
a sum-all-reduce of the flat fp32 gradient arena (NCCL over NVLink on GPUs; gloo in the CPU tests),
scaled by 1/world inside the Adamax launch.  BatchNorm statistics, KL-balance coefficients and
the batch mean of the loss stay per replica, as in the (single-replica) reference.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> int:
    """torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend)
    return world


def bucket_bounds(n: int, bucket_elems: int) -> List[slice]:
    """Contiguous buckets over the flat arena, last-to-first (backward produces the tail of the
    arena -- postprocess -- first, so those buckets can be reduced while the rest is still computing)."""
    if bucket_elems <= 0 or bucket_elems >= n:
        return [slice(0, n)]
    bounds = []
    hi = n
    while hi > 0:
        lo = max(0, hi - bucket_elems)
        bounds.append(slice(lo, hi))
        hi = lo
    return bounds


def all_reduce_gradients(flat: torch.Tensor, group=None, bucket_elems: int = 0, async_op: bool = False):
    """Sum `flat` over the ranks of `group` in place.  Returns the list of work handles when async."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return []
    works = []
    for sl in bucket_bounds(flat.numel(), bucket_elems):
        w = dist.all_reduce(flat[sl], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def broadcast_parameters(tensors: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Rank `src`'s parameter/state arenas become everybody's (identical replicas at step 0)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src, group=group)


def shard_batch(global_batch: int, rank: int, world: int) -> slice:
    """Samples [lo, hi) of a global batch owned by `rank` (equal shards; the global batch must divide)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} does not divide over {world} ranks")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)
