"""WGAN-GP train step timing (WassersteinGAN.py defaults: 64x64 masks, batch 64, latent 128, 3 critic updates + 1 generator
update per step): python scripts/bench_wgan.py [--size 64] [--batch 64] [--dtype bf16] [--steps 20]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sem_b200
from sem_b200 import WganGpModel, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
m = WganGpModel((a.size, a.size, 1), batch_size=a.batch, latent_dim=128, dtype=a.dtype)
x = (np.random.default_rng(0).random((a.batch, a.size, a.size, 1)) > 0.7).astype(np.float32) * 2 - 1
c0 = _lib.launch_count()
logs = m.train_step(x)
per_step = _lib.launch_count() - c0
for _ in range(3):
    m.train_step(x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(a.steps):
    logs = m.train_step(x)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / a.steps
print(json.dumps({"workload": f"WGAN-GP train step {a.size}x{a.size} batch {a.batch} latent 128 {a.dtype} (3 critic + 1 generator update, eager)",
                  "ms_per_step": ms, "masks_per_s": a.batch / ms * 1e3, "launches_per_step": per_step, "metrics": logs}))
