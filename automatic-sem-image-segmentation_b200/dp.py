"""Data-parallel plumbing: one process per GPU, tiles sharded by rank, ONE all-reduce of the flat gradient buffer per
network per step (torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference has no working multi-GPU path on the torch backend (keras.distribution.DataParallel is JAX-only in
Keras 3.5); this is new functionality with DDP semantics: per-rank BatchNorm statistics,
gradients averaged over ranks (sum here, 1/world folded into the fused Adam's gscale).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """rank-strided slice of a common (already shuffled) index list; every rank gets floor(n/world) items."""
    per = n_items // world
    return [rank + i * world for i in range(per)]


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks of one flat buffer (latency-bound for the UNet's 10 MB: a single collective)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_(flat: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat
