"""The product's data-parallel path on the ENGINE (not on oracle gradients): two processes share cuda:0, each holds a
UNetModel replica, `set_distributed()` + `train_step()` run the real sequence (fwd+bwd -> flat all-reduce of the engine's
gradient buffer -> Adam with 1/world folded in).  torch.distributed runs over gloo here (NCCL refuses two ranks on one
device; gloo all-reduces CUDA tensors through the host), so the collective call sites, buffer and scaling are the
product's, only the transport differs from the 8-GPU runs.

(This test found a real race: the first upload of a training instance ran on the copy stream before the zero fill of its
freshly allocated staging buffers had executed on the main stream -- an all-zero first batch, once in three runs on a GPU
shared by two processes; see model._TrainIO.)

Checked on rank 0: (1) rank 1's weights equal rank 0's after the broadcast although they were initialised differently,
(2) the engine's gradient buffer after the step is the SUM of the two ranks' local gradients (each recomputed by a
world_size-1 replica on the same half batch), (3) both ranks hold identical weights after Adam, and they equal a
single-process Adam step (gscale = 1/world) on that buffer."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import sem_b200  # noqa: F401
    from sem_b200 import UNetModel, dp
    from oracle import unet as OU
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, h, w = 4, 32, 32
        x, y, wgt = OU.synthetic_batch(n, h, w)
        idx = dp.shard_indices(n, rank, world)
        xs, ys = x[idx].numpy(), y[idx].numpy()
        spec = OU.UNetSpec(16)
        p_rank = {k: v.detach().numpy() for k, v in spec.init_params(seed=rank).items()}      # different weights per rank
        p0 = {k: v.detach().numpy() for k, v in spec.init_params(seed=0).items()}

        # local gradient of this rank's half batch with rank 0's weights, from a replica that knows nothing of the group
        solo = UNetModel((h, w, 1), 16, dtype="f32", batch_size=len(idx), use_cuda_graph=False)
        solo.set_named_weights(p0)
        solo.compile(weighting=wgt)
        inst = solo._use(solo._instance(len(idx), h, w))
        inst.x_dev.copy_(torch.from_numpy(xs)); inst.y_dev.copy_(torch.from_numpy(ys))
        inst.stage_in()
        inst.fwd_bwd(wgt)
        torch.cuda.synchronize()
        g_local = inst.eng.grads.detach().clone()
        w_before = inst.eng.params.t.detach().clone()

        m = UNetModel((h, w, 1), 16, dtype="f32", batch_size=len(idx), use_cuda_graph=False)
        m.set_named_weights(p_rank)
        m.compile(weighting=wgt)
        m._use(m._instance(len(idx), h, w))
        m.set_distributed()
        torch.cuda.synchronize()
        e = m.engine
        bcast_err = float((e.params.t - w_before).abs().max())          # (1) rank 0's weights everywhere
        seen = {}
        orig_allreduce = m._allreduce

        def spy(eng):                                                   # the engine's LOCAL gradient right before the collective
            seen["local"] = eng.grads.detach().clone()
            orig_allreduce(eng)

        m._allreduce = spy
        logs = m.train_step(xs, ys)
        torch.cuda.synchronize()
        local_err = float((seen["local"] - g_local).abs().max() / g_local.abs().max())
        if local_err > 1e-4:                                            # which tensors differ (diagnostics for the assertion message)
            bad = []
            for name, (o, cnt) in e.params.entries.items():
                a, b = seen["local"][o:o + cnt], g_local[o:o + cnt]
                d = float((a - b).abs().max())
                if d > 1e-6 * float(g_local.abs().max()):
                    bad.append((name, float(a.abs().max()), float(b.abs().max()), d))
            ret[f"bad{rank}"] = bad[:12] + [("count", len(bad), len(e.params.entries), 0.0)]
        g_sum = g_local.clone()
        dist.all_reduce(g_sum)                                          # reference: sum of the two local gradients
        grad_err = float((e.grads - g_sum).abs().max() / g_sum.abs().max())
        w_after = e.params.t.detach().clone()
        # single-process Adam fed with the product's all-reduced buffer (the solo replica: same weights, zero Adam state);
        # a first Adam step is lr * sign(g), so it must see the very same buffer, not a re-summed one
        inst.eng.grads.copy_(e.grads)
        inst.eng.lr.fill_(solo.learning_rate)
        inst.eng.adam(solo.beta_1, solo.beta_2, solo.epsilon, 1.0 / world)
        torch.cuda.synchronize()
        adam_err = float((inst.eng.params.t - w_after).abs().max())
        w_all = [torch.zeros_like(w_after) for _ in range(world)]
        dist.all_gather(w_all, w_after)
        if rank == 0:
            ret["bcast_err"] = bcast_err
            ret["grad_err"] = grad_err
            ret["local_err"] = local_err
            ret["adam_err"] = adam_err
            ret["rank_diff"] = float((w_all[0] - w_all[1]).abs().max())
            ret["moved"] = float((w_after - w_before).abs().max())
            ret["loss"] = float(logs["loss"])
        else:
            ret["bcast_err1"] = bcast_err
            ret["local_err1"] = local_err
    finally:
        dist.destroy_process_group()


def test_two_ranks_on_one_gpu_run_the_engine_dp_path():
    ret = mp.Manager().dict()
    port = 29500 + (os.getpid() * 7) % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["bcast_err"] == 0.0 and ret["bcast_err1"] == 0.0, dict(ret)
    assert ret["local_err"] < 1e-4 and ret["local_err1"] < 1e-4, dict(ret)
    assert ret["grad_err"] < 1e-4, dict(ret)          # fp32 atomics order inside one rank's backward: run-to-run noise only
    assert ret["rank_diff"] == 0.0, dict(ret)
    assert ret["adam_err"] < 1e-6 and ret["moved"] > 1e-4, dict(ret)
    assert np.isfinite(ret["loss"])
