"""world_size-2 gloo test of the data-parallel path: gradients of two ranks on half batches, summed with the
product's flat all-reduce and scaled by 1/world, equal the single-process gradient on the concatenated batch.
(InstanceNorm network -> exact; BatchNorm statistics are per rank by design, see DESIGN.md.)"""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sem_b200
from sem_b200 import dp
from oracle import cyclegan as OC, layers as OL


def _flat_grads(params, x):
    for v in params.values():
        v.grad = None
    out = OC.discriminator_forward(x, params)
    OL.mse(torch.ones_like(out), out).backward()
    return torch.cat([v.grad.reshape(-1) for v in params.values()])


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(0)
    params = OC.init_params(OC.discriminator_spec(8), gen)
    flat_p = torch.cat([v.reshape(-1) for v in params.values()])
    if rank != 0:
        flat_p.zero_()
    dp.broadcast_(flat_p, 0)                       # rank 0's weights everywhere
    off = 0
    for v in params.values():
        v.data.copy_(flat_p[off:off + v.numel()].reshape(v.shape)); off += v.numel()
        v.requires_grad_(True)
    x = torch.rand(4, 64, 64, 1, generator=torch.Generator().manual_seed(1)) * 2 - 1
    idx = dp.shard_indices(4, rank, world)
    g = _flat_grads(params, x[idx])
    dp.allreduce_sum_(g)
    g /= world
    if rank == 0:
        full = _flat_grads(params, x)
        ret["err"] = float((g - full).abs().max() / full.abs().max())
        ret["idx"] = idx
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_rank():
    assert dp.shard_indices(10, 1, 4) == [1, 5] and dp.shard_indices(4, 0, 2) == [0, 2]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["idx"] == [0, 2]
    assert ret["err"] < 1e-5, ret["err"]


def test_inference_work_list_is_rank_strided(monkeypatch):
    """run_inference shards the file list by rank under torchrun (no collective) and keeps every file in a
    single process."""
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    assert dp.inference_indices(5) == [0, 1, 2, 3, 4]
    monkeypatch.setenv("WORLD_SIZE", "4")
    got = []
    for r in range(4):
        monkeypatch.setenv("RANK", str(r))
        got.append(dp.inference_indices(10))
    assert got == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert sorted(i for g in got for i in g) == list(range(10))
