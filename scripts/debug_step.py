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
tr = OU.UNetTrainer(spec, p0, wgt)
logs_ref, yp_ref = tr.train_step(x, y)
gmax = max(float(v.abs().max()) for v in tr.last_grads.values())
print("oracle", logs_ref, "gmax", gmax)
runs = []
for rep in range(3):
    m = UNetModel((h, w, 1), 16, dtype="f32", batch_size=n, use_cuda_graph=False)
    m.set_named_weights({k: v.detach().numpy() for k, v in p0.items()})
    m.compile(weighting=wgt)
    logs = m.train_step(x.numpy(), y.numpy())
    e = m.engine
    gr = {name: e.get_grad(name) for name in spec.trainable_names()}
    runs.append(gr)
    print("run", rep, logs)
for name in spec.trainable_names():
    ref = tr.last_grads[name].numpy()
    errs = [np.abs(r[name] - ref).max() for r in runs]
    var = np.abs(runs[0][name] - runs[1][name]).max()
    flag = "  <<<" if max(errs) > 1e-3 * max(np.abs(ref).max(), 1e-3 * gmax) else ""
    print(f"{name:40s} refmax {np.abs(ref).max():.3e} err {errs[0]:.3e} {errs[1]:.3e} {errs[2]:.3e} run2run {var:.3e}{flag}")
