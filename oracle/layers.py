"""Keras-3.5 torch-backend layer semantics restated in plain torch (CPU).

TEST INFRASTRUCTURE (see oracle/__init__.py).  All tensors are NHWC like the
reference's Keras tensors; every conv/pool permutes to NCHW, calls
torch.nn.functional and permutes back, which is what
keras/src/backend/torch/nn.py does [K3.5, SURVEY.md Appendix B].

Weight layouts are Keras': Conv2D kernel HWIO (kh,kw,Cin,Cout); Conv2DTranspose
kernel (kh,kw,Cout,Cin); BatchNormalization [gamma], beta, moving_mean,
moving_variance; GroupNormalization gamma, beta.
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.99   # keras.layers.BatchNormalization default (UNet_Segmentation.py:422)
BN_EPS = 1e-3        # keras default epsilon
IN_EPS = 1e-5        # CycleGAN.py:329 epsilon=1e-5
BCE_CLIP = 1e-7      # keras.backend.epsilon()


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def same_pad_amounts(size: int, k: int, stride: int):
    """Keras torch backend 'same' padding for stride>1 (Appendix B item 2):
    total = (k-1) - (size-1) % stride ; left = total//2 ; right = (total+1)//2."""
    total = max((k - 1) - (size - 1) % stride, 0)
    return total // 2, (total + 1) // 2


def conv2d(x, kernel, bias=None, stride: int = 1, padding: str = "same"):
    """keras.layers.Conv2D(...)(x).  x NHWC, kernel HWIO.
    Reference call sites: UNet_Segmentation.py:421, CycleGAN.py:327,340,372,429."""
    kh, kw = kernel.shape[0], kernel.shape[1]
    w = kernel.permute(3, 2, 0, 1).contiguous()  # OIHW (contiguous: the fp64 CPU conv backward requires it)
    xi = _nchw(x)
    if padding == "same":
        if stride == 1:
            y = F.conv2d(xi, w, bias, stride=1, padding="same")
        else:
            pt, pb = same_pad_amounts(x.shape[1], kh, stride)
            pl, pr = same_pad_amounts(x.shape[2], kw, stride)
            xi = F.pad(xi, (pl, pr, pt, pb))
            y = F.conv2d(xi, w, bias, stride=stride, padding=0)
    elif padding == "valid":
        y = F.conv2d(xi, w, bias, stride=stride, padding=0)
    else:
        raise ValueError(padding)
    return _nhwc(y)


def conv_transpose_pads(k: int, stride: int):
    """Keras -> torch padding / output_padding for Conv2DTranspose 'same',
    output_padding=None (Appendix B item 3)."""
    out_pad_keras = stride - k % 2
    # keras: _convert_conv_transpose_padding_args_from_keras_to_torch
    torch_padding = max(-((k % 2 - k + out_pad_keras) // 2), 0)
    torch_output_padding = 2 * torch_padding + k % 2 - k + out_pad_keras
    return torch_padding, torch_output_padding


def conv2d_transpose(x, kernel, bias=None, stride: int = 2):
    """keras.layers.Conv2DTranspose(filters,(k,k),strides=stride,padding='same').
    kernel (kh,kw,Cout,Cin).  UNet_Segmentation.py:542-551, CycleGAN.py:353."""
    k = kernel.shape[0]
    p, op = conv_transpose_pads(k, stride)
    w = kernel.permute(3, 2, 0, 1).contiguous()  # (Cin, Cout, kh, kw)
    y = F.conv_transpose2d(_nchw(x), w, bias, stride=stride, padding=p, output_padding=op)
    return _nhwc(y)


def batch_norm(x, gamma, beta, moving_mean, moving_var, training: bool,
               momentum: float = BN_MOMENTUM, eps: float = BN_EPS):
    """keras.layers.BatchNormalization(axis=3).  Returns (y, new_mean, new_var).
    Training: mean=E[x], var=E[x^2]-E[x]^2 (biased) over N,H,W; moving stats are
    updated with the *biased* variance (Appendix B item 4)."""
    if training:
        mean = x.mean(dim=(0, 1, 2))
        var = (x * x).mean(dim=(0, 1, 2)) - mean * mean
        new_mean = moving_mean * momentum + mean.detach() * (1.0 - momentum)
        new_var = moving_var * momentum + var.detach() * (1.0 - momentum)
    else:
        mean, var = moving_mean, moving_var
        new_mean, new_var = moving_mean, moving_var
    inv = torch.rsqrt(var + eps)
    if gamma is not None:
        inv = inv * gamma
    y = (x - mean) * inv + beta
    return y, new_mean, new_var


def instance_norm(x, gamma, beta, eps: float = IN_EPS):
    """keras.layers.GroupNormalization(groups=-1, axis=3): per-(n,c) statistics over
    H,W, var = E[x^2]-E[x]^2 (Appendix B item 5).  CycleGAN.py:329,335,342,355,374."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = (x * x).mean(dim=(1, 2), keepdim=True) - mean * mean
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def reflection_pad(x, pad_w_total: int, pad_h_total: int):
    """ReflectionPadding2D((w_total,h_total)) -- UNet_Segmentation.py:578-589.
    Totals are split total//2 before and total//2 + total%2 after."""
    if pad_w_total == 0 and pad_h_total == 0:
        return x
    pt, pb = pad_h_total // 2, pad_h_total // 2 + pad_h_total % 2
    pl, pr = pad_w_total // 2, pad_w_total // 2 + pad_w_total % 2
    return _nhwc(F.pad(_nchw(x), (pl, pr, pt, pb), mode="reflect"))


def max_pool_2x2(x):
    """keras.layers.MaxPooling2D((2,2)) -- stride 2, valid."""
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))


def leaky_relu(x, slope: float = 0.2):
    return F.leaky_relu(x, slope)


def weighted_bce(y_true, y_pred, weighting: float):
    """UNet_Segmentation.py:379-384.  keras BinaryCrossentropy(reduction='none')
    clips p to [1e-7, 1-1e-7] and averages over the last axis."""
    p = torch.clamp(y_pred, BCE_CLIP, 1.0 - BCE_CLIP)
    bce = -(y_true * torch.log(p) + (1.0 - y_true) * torch.log(1.0 - p))
    bce = bce.mean(dim=-1, keepdim=True)
    weights = y_true * (weighting - 1.0) + 1.0
    return (bce * weights).mean()


def mse(a, b):
    return ((a - b) ** 2).mean()


def mae(a, b):
    return (a - b).abs().mean()


class KerasAdam:
    """keras.optimizers.Adam (Appendix B item 8): epsilon outside the bias-corrected
    sqrt: alpha_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= alpha_t*m/(sqrt(v)+eps)."""

    def __init__(self, params, lr=1e-3, beta_1=0.9, beta_2=0.999, eps=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, eps
        self.t = 0
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]

    @torch.no_grad()
    def apply(self, grads, params):
        self.t += 1
        t = self.t
        alpha = self.lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        for p, g, m, v in zip(params, grads, self.m, self.v):
            m.add_((g - m) * (1.0 - self.b1))
            v.add_((g * g - v) * (1.0 - self.b2))
            p.sub_(alpha * m / (v.sqrt() + self.eps))


def glorot_uniform(shape, gen: torch.Generator, fan_in: int, fan_out: int):
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2.0 - 1.0) * limit


# ---- storage-precision emulation (TEST INFRASTRUCTURE) ---------------------------------------------------------------
# The CUDA path's throughput mode STORES activations and their gradients as bf16 (fp32 accumulation inside every kernel).
# `storage(torch.bfloat16)` makes the oracle round at the same tensor boundaries (forward values and backward gradients),
# which measures how far ANY bf16-storage implementation must drift from the fp32 path on a given problem; the GPU parity
# tests state their per-tensor tolerance as a multiple of that drift.
_STORAGE = None


class storage:
    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _STORAGE
        self.prev, _STORAGE = _STORAGE, self.dtype

    def __exit__(self, *a):
        global _STORAGE
        _STORAGE = self.prev


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.dtype = dtype
        return x.to(dtype).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dtype).to(g.dtype), None


def st(x):
    """A tensor boundary at which the CUDA path writes to HBM: identity unless a storage dtype is being emulated."""
    return x if _STORAGE is None else _RoundSTE.apply(x, _STORAGE)
