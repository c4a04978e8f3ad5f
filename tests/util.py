"""Helpers shared by the GPU parity tests: torch tensors <-> C-ABI views, channel padding."""
import ctypes as C

import numpy as np
import torch

import sem_b200  # noqa: F401
from sem_b200 import _lib as L


def pad8(c):
    return (c + 7) // 8 * 8


def tdtype(dtype):
    return torch.bfloat16 if dtype == "bf16" else torch.float32


def ldtype(dtype):
    return L.BF16 if dtype == "bf16" else L.F32


def to_dev(x_nhwc: torch.Tensor, dtype: str, pitch=None, coff=0):
    """logical NHWC fp32 (cpu) -> padded device tensor (N,H,W,pitch) with the data at channels [coff, coff+C)."""
    n, h, w, c = x_nhwc.shape
    pitch = pitch or pad8(c)
    t = torch.zeros((n, h, w, pitch), dtype=tdtype(dtype), device="cuda")
    t[..., coff:coff + c] = x_nhwc.to("cuda").to(tdtype(dtype))
    return t


def view(t: torch.Tensor, coff=0, c=None):
    c = t.shape[-1] - coff if c is None else c
    return L.Tensor(t.data_ptr(), c, t.shape[-1], coff)


def pad_w(w_hwio: torch.Tensor):
    """fp32 HWIO logical -> physical (both channel axes padded to 8), on device."""
    r, s, ci, co = w_hwio.shape
    out = torch.zeros((r, s, pad8(ci), pad8(co)), dtype=torch.float32)
    out[:, :, :ci, :co] = w_hwio
    return out.cuda().contiguous()


def pad_v(v: torch.Tensor):
    out = torch.zeros(pad8(v.shape[0]), dtype=torch.float32)
    out[:v.shape[0]] = v
    return out.cuda()


def stream():
    return torch.cuda.current_stream().cuda_stream


def rel_err(a: torch.Tensor, b: torch.Tensor):
    """max |a-b| normalised by max |b| (SURVEY.md 8c tolerance definition)."""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def bf16_round(x: torch.Tensor):
    return x.to(torch.bfloat16).to(torch.float32)
