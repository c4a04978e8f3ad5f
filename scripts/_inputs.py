"""Synthetic inputs of the bench workload for the profiling scripts (BASELINE config 2): x ~ U[0,1) seed 0,
y = (U[0,1) seed 1 < 0.10), weighting = #zeros / #ones -- the same recipe bench.py uses, restated here so that nothing under
scripts/ touches oracle/ (the oracle is the checker of tests/, smoke() and bench.py's CPU arm only)."""
import torch


def synthetic_batch(n: int, h: int = 256, w: int = 256):
    x = torch.rand(n, h, w, 1, generator=torch.Generator().manual_seed(0))
    y = (torch.rand(n, h, w, 1, generator=torch.Generator().manual_seed(1)) < 0.10).to(torch.float32)
    ones = float(y.sum().item())
    return x, y, float(y.numel() - ones) / max(ones, 1.0)
