"""Loss trajectories of eager vs CUDA-graph training steps (f32 and bf16) under the current env switches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sem_b200
from sem_b200 import UNetModel
from oracle import unet as OU
n, h, w = 2, 32, 32
x, y, wgt = OU.synthetic_batch(n, h, w)
for dtype in ("f32", "bf16"):
    out = []
    for graph in (False, False, True, True):
        m = UNetModel((h, w, 1), 16, dtype=dtype, batch_size=n, seed=3, use_cuda_graph=graph)
        m.compile(weighting=wgt)
        out.append([m.train_step(x.numpy(), y.numpy())["loss"] for _ in range(4)])
    print(dtype, {k: os.environ.get(k) for k in ("SEMB_NO_WGRAD_STREAM", "SEMB_NO_FOLD_NORM")})
    for o in out:
        print("   ", ["%.7f" % v for v in o])
