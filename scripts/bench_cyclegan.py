"""CycleGAN train step (BASELINE.json configs[2]: G_A2B/G_B2A + PatchGAN-D, 256x256, batch 8, image pool) on one B200.

    python scripts/bench_cyclegan.py [--filters 64] [--batch 8] [--steps 5]

Reports image pairs / s and the achieved conv TFLOP/s against the algorithmic minimum of SURVEY 8d (1967 GFLOP per image
pair at filters=64, 25.78/102.29 of it at filters=32 for the generators)."""
import argparse, json, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sem_b200
from sem_b200 import CycleGanModel, ImagePool

ap = argparse.ArgumentParser()
ap.add_argument("--filters", type=int, default=64)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--dtype", default="bf16")
args = ap.parse_args()
n, sz = args.batch, args.size
m = CycleGanModel((sz, sz, 1), batch_size=n, filters=args.filters, dtype=args.dtype,
                  image_pool_a=ImagePool(n, 50, random.Random(0)), image_pool_b=ImagePool(n, 50, random.Random(1)))
m.compile()
for net in m.nets.values():          # Glorot weights (the constructor leaves them zero until set_weights)
    net.root.init_params(seed=0)
g = torch.Generator().manual_seed(0)
a = (torch.rand(n, sz, sz, 1, generator=g) * 2 - 1).numpy()
b = (torch.rand(n, sz, sz, 1, generator=g) * 2 - 1).numpy()
for _ in range(2):
    logs = m.train_step((a, b))
torch.cuda.synchronize()
if os.environ.get("PROFILE"):          # ncu --profile-from-start off: one profiled step
    torch.cuda.cudart().cudaProfilerStart()
    m.train_step((a, b))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(args.steps):
    logs = m.train_step((a, b))
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / args.steps
gflop_pair = 1967.0 * (args.filters / 64.0) ** 2 * (sz / 256.0) ** 2
print(json.dumps({"workload": f"CycleGAN train step {sz}x{sz} batch {n} filters {args.filters} {args.dtype}", "ms_per_step": ms,
                  "image_pairs_per_s": n / ms * 1e3, "conv_TFLOPs_algorithmic_min": gflop_pair * n / ms,
                  "metrics": {k: round(float(v), 5) for k, v in logs.items()}}))
