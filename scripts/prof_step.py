"""Two eager train steps at the bench workload, for ncu (launch list / full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sem_b200
from sem_b200 import UNetModel, _lib
from oracle import unet as OU
n = int(os.environ.get("BATCH", "32"))
x, y, wgt = OU.synthetic_batch(n, 256, 256)
m = UNetModel((256, 256, 1), 16, dtype="bf16", batch_size=n, use_cuda_graph=False)
m.compile(weighting=wgt)
c0 = _lib.launch_count()
for i in range(int(os.environ.get("STEPS", "2"))):
    logs = m.train_step(x.numpy(), y.numpy())
    print("step", i, logs, "launches so far", _lib.launch_count() - c0, flush=True)
