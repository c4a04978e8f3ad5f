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


def inference_indices(n_items: int) -> List[int]:
    """The items THIS process handles in run_inference: all of them in a single process; under torchrun (RANK / WORLD_SIZE
    in the environment) a rank-strided slice of the file list -- images are independent, every rank writes its own output
    files, there is no collective.  Unlike shard_indices the tail is not dropped: rank r gets r, r+W, r+2W, ..."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1") or 1)
    rank = int(os.environ.get("RANK", "0") or 0)
    if dist.is_initialized():
        world, rank = dist.get_world_size(), dist.get_rank()
    if world <= 1:
        return list(range(n_items))
    return list(range(rank, n_items, world))
