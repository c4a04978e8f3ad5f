import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import unet as OU
import sem_b200
from sem_b200 import UNetModel

n, h, w = 2, 32, 32
x, y, wgt = OU.synthetic_batch(n, h, w)
models = []
snaps = []
for rep in range(6):
    m = UNetModel((h, w, 1), 16, dtype="f32", batch_size=n, seed=0, use_cuda_graph=False)
    m.compile(weighting=wgt)
    inst = m._current
    inst.x_dev.copy_(x.cuda()); inst.y_dev.copy_(y.cuda())
    inst.fwd_bwd(wgt)
    torch.cuda.synchronize()
    e = inst.eng
    snap = {}
    for b in e.bufs:
        snap["D:" + b.name] = b.data.float().cpu().clone()
        if b._grad is not None:
            snap["G:" + b.name] = b._grad.float().cpu().clone()
    snap["Z:zeroed"] = e.zeroed.t.cpu().clone()
    snap["S:scratch"] = e.scratch.t.cpu().clone()
    snap["P:grads"] = e.grads.cpu().clone()
    snaps.append(snap)
    models.append(m if rep % 2 == 0 else None)   # vary memory reuse
ref = snaps[0]
for rep in range(1, 6):
    bad = []
    for k in ref:
        d = (snaps[rep][k] - ref[k]).abs().max().item()
        s = ref[k].abs().max().item()
        if d > 1e-4 * max(s, 1e-6):
            bad.append((k, d, s))
    print("rep", rep, "n_bad", len(bad), bad[:12])
# localise inside zeroed/scratch for the worst rep
e = models[0]._current.eng
for rep in range(1, 6):
    for store, key in ((e.zeroed, "Z:zeroed"), (e.scratch, "S:scratch")):
        for name in store.order:
            o, nn = store.entries[name]
            a, b = ref[key][o:o+nn], snaps[rep][key][o:o+nn]
            d = (a - b).abs().max().item(); s = a.abs().max().item()
            if d > 1e-4 * max(s, 1e-6):
                print("   rep", rep, name, d, s)
                break
