"""Experiment: per-tensor gradient agreement GPU vs oracle at 256x256 batch 32 for random vs real data, bf16 vs f32 storage."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import unet as OU
from sem_b200 import UNetModel

def bf16r(t): return t.to(torch.bfloat16).float()

def run(kind, dtype, n=32):
    spec = OU.UNetSpec(16)
    p0 = {k: bf16r(v) if k.endswith("/kernel") else v for k, v in spec.init_params(seed=0).items()}
    if kind == "random":
        x, y, wgt = OU.synthetic_batch(n, 256, 256)
    else:
        with np.load("baseline/_ref/sem_dataset.npz") as z:
            imgs, masks = z["images"], np.unpackbits(z["masks"], axis=-1).astype(bool)
        xs, ys = [], []
        rng = np.random.default_rng(0)
        for i in range(n):
            r, c = int(rng.integers(0, 704 - 256)), int(rng.integers(0, 1024 - 256))
            a = imgs[i % 40][r:r + 256, c:c + 256].astype(np.float32)
            xs.append((a - a.min()) / max(float((a - a.min()).max()), 1.0)); ys.append(masks[i % 40][r:r + 256, c:c + 256].astype(np.float32))
        x, y = torch.from_numpy(np.stack(xs))[..., None], torch.from_numpy(np.stack(ys))[..., None]
        wgt = float((y == 0).sum() / (y == 1).sum())
    x = bf16r(x)
    torch.set_num_threads(os.cpu_count())
    tr = OU.UNetTrainer(spec, p0, wgt)
    ref, _ = tr.train_step(x, y)
    m = UNetModel((256, 256, 1), 16, dtype=dtype, batch_size=n, use_cuda_graph=False)
    m.set_named_weights({k: v.numpy() for k, v in p0.items()})
    m.compile(weighting=wgt)
    logs = dict(m.train_step(x.numpy(), y.numpy()))
    e = m.engine
    gmax = max(float(g.abs().max()) for g in tr.last_grads.values())
    cos, err = [], []
    for name in spec.trainable_names():
        r = tr.last_grads[name].float(); g = torch.from_numpy(e.get_grad(name))
        if float(r.abs().max()) < 1e-3 * gmax: continue
        cos.append((float((r * g).sum() / (r.norm() * g.norm()).clamp_min(1e-30)), name))
        err.append((float((r - g).abs().max() / r.abs().max()), name))
    cos.sort(); err.sort()
    print(kind, dtype, "loss", logs["loss"], ref["loss"], "n", len(cos), "cos min/med", cos[0], cos[len(cos)//2][0], "err max/med", err[-1], err[len(err)//2][0], flush=True)

for kind in ("real", "random"):
    for dtype in ("bf16", "f32"):
        run(kind, dtype, int(os.environ.get("N", "32")))
