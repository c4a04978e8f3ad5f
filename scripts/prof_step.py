"""Eager train steps at the bench workload for ncu (launch lists / metric passes / full captures).

    STEPS=1 WARM=1 ncu --profile-from-start off ... python scripts/prof_step.py

WARM un-profiled steps first (lazy gradient-buffer allocation and its fill kernels), then cudaProfilerStart and STEPS
profiled steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import sem_b200
from sem_b200 import UNetModel, _lib
from _inputs import synthetic_batch          # (scripts/ does not import oracle/)
n = int(os.environ.get("BATCH", "32"))
x, y, wgt = synthetic_batch(n, 256, 256)
m = UNetModel((256, 256, 1), 16, dtype="bf16", batch_size=n, use_cuda_graph=False)
m.compile(weighting=wgt)
for i in range(int(os.environ.get("WARM", "1"))):
    m.train_step(x.numpy(), y.numpy())
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
c0 = _lib.launch_count()
for i in range(int(os.environ.get("STEPS", "1"))):
    logs = m.train_step(x.numpy(), y.numpy())
    print("step", i, logs, "launches so far", _lib.launch_count() - c0, flush=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
