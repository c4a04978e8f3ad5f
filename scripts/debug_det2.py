import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import unet as OU
import sem_b200
from sem_b200 import UNetModel

n, h, w = 2, 32, 32
spec = OU.UNetSpec(16)
p0 = spec.init_params(seed=0)
g = torch.Generator().manual_seed(123)
x = torch.rand(n, h, w, 1, generator=g)
y = (torch.rand(n, h, w, 1, generator=g) < 0.2).float()
wgt = float((y == 0).sum() / (y == 1).sum())
named = {k: v.detach().numpy() for k, v in p0.items()}
snaps = []
keep = []
for rep in range(40):
    m = UNetModel((h, w, 1), 16, dtype="f32", batch_size=n, use_cuda_graph=False)
    m.set_named_weights(named)
    m.compile(weighting=wgt)
    logs = m.train_step(x.numpy(), y.numpy())
    torch.cuda.synchronize()
    e = m._current.eng
    snap = {}
    order = []
    for b in e.bufs:
        snap["D:" + b.name] = b.data.float().cpu().clone(); order.append("D:" + b.name)
    for b in reversed(e.bufs):
        if b._grad is not None:
            snap["G:" + b.name] = b._grad.float().cpu().clone(); order.append("G:" + b.name)
    snap["S:scratch"] = e.scratch.t.cpu().clone(); order.append("S:scratch")
    snap["P:grads"] = e.grads.cpu().clone(); order.append("P:grads")
    snaps.append(snap)
    if rep % 3 == 0: keep.append(m)
ref = snaps[0]
for rep in range(1, len(snaps)):
    bad = []
    for k in order:
        d = (snaps[rep][k] - ref[k]).abs().max().item()
        s = ref[k].abs().max().item()
        if d > 1e-4 * max(s, 1e-6):
            bad.append((k, round(d / max(s, 1e-6), 6)))
    print("rep", rep, "n_bad", len(bad), bad[:6])

for rep in range(1, len(snaps)):
    a, b = ref["G:mres9_cat_act"], snaps[rep]["G:mres9_cat_act"]
    d = (a - b).abs()
    idx = (d > 1e-3 * a.abs().max()).nonzero()
    if len(idx):
        print("rep", rep, "n differing elems in G:mres9_cat_act:", len(idx), idx[:5].tolist())
        for ii in idx[:5]:
            n_, y_, x_, c_ = ii.tolist()
            print("    out1 ref/rep:", ref["D:mres9_sum"][n_, y_, x_, c_].item(), snaps[rep]["D:mres9_sum"][n_, y_, x_, c_].item(),
                  "grad ref/rep", a[n_, y_, x_, c_].item(), b[n_, y_, x_, c_].item())
