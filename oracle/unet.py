"""MultiRes-UNet of the reference restated in plain torch (CPU).  TEST INFRASTRUCTURE.

Follows /root/reference/Releases/Version 1.2.0/UNet_Segmentation.py:
  conv2d_bn        :401-426   Conv2D(no bias) -> BatchNormalization(scale=False) -> act
  multi_res_block  :451-474
  res_path         :476-503
  multi_res_unet   :505-562
  weighted_bce / Adam / compile :379-395
  TorchTrainer.train_step [K3.5]: fwd(training=True) -> loss -> backward -> Adam

Parameters are kept in *creation (call) order*, which is also the numbering of
the Const nodes in the reference's frozen graphs (conv2d_{1..57}_1,
conv2d_transpose_{1..4}_1, batch_normalization_{1..85}_1).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import layers as L


def mres_widths(u: int, alpha: float = 1.67):
    w = alpha * u
    return int(w * 0.167), int(w * 0.333), int(w * 0.5)


class UNetSpec:
    """Walks multi_res_unet once, recording every variable in creation order."""

    def __init__(self, filters: int = 16, in_channels: int = 1, output_channels: int = 1):
        self.filters = filters
        self.in_channels = in_channels
        self.output_channels = output_channels
        self.entries = []  # (name, shape, kind)
        self._nconv = 0
        self._nbn = 0
        self._nct = 0
        self._build()

    # -- recording helpers
    def _conv(self, cin, cout, k):
        self._nconv += 1
        self.entries.append((f"conv2d_{self._nconv}/kernel", (k, k, cin, cout), "conv_kernel"))

    def _bn(self, c, scale):
        self._nbn += 1
        n = f"batch_normalization_{self._nbn}"
        if scale:
            self.entries.append((n + "/gamma", (c,), "bn_gamma"))
        self.entries.append((n + "/beta", (c,), "bn_beta"))
        self.entries.append((n + "/moving_mean", (c,), "bn_mean"))
        self.entries.append((n + "/moving_variance", (c,), "bn_var"))

    def _convT(self, cin, cout):
        self._nct += 1
        n = f"conv2d_transpose_{self._nct}"
        self.entries.append((n + "/kernel", (2, 2, cout, cin), "convT_kernel"))
        self.entries.append((n + "/bias", (cout,), "convT_bias"))

    def _conv_bn(self, cin, cout, k):
        self._conv(cin, cout, k)
        self._bn(cout, scale=False)

    def _mres(self, u, cin):
        a, b, c = mres_widths(u)
        tot = a + b + c
        self._conv_bn(cin, tot, 1)
        self._conv_bn(cin, a, 3)
        self._conv_bn(a, b, 3)
        self._conv_bn(b, c, 3)
        self._bn(tot, True)
        self._bn(tot, True)
        return tot

    def _respath(self, f, length, cin):
        for _ in range(length):
            self._conv_bn(cin, f, 1)
            self._conv_bn(cin, f, 3)
            self._bn(f, True)
            cin = f
        return f

    def _build(self):
        f = self.filters
        c1 = self._mres(f, self.in_channels)
        self._respath(f, 4, c1)
        c2 = self._mres(f * 2, c1)
        self._respath(f * 2, 3, c2)
        c3 = self._mres(f * 4, c2)
        self._respath(f * 4, 2, c3)
        c4 = self._mres(f * 8, c3)
        self._respath(f * 8, 1, c4)
        c5 = self._mres(f * 16, c4)
        self._convT(c5, f * 8)
        c6 = self._mres(32 * 8, f * 16)       # decoder widths hard-coded 32*k (:543-549)
        self._convT(c6, f * 4)
        c7 = self._mres(32 * 4, f * 8)
        self._convT(c7, f * 2)
        c8 = self._mres(32 * 2, f * 4)
        self._convT(c8, f)
        c9 = self._mres(f, f * 2)
        if self.output_channels == 1:
            self._conv_bn(c9, 1, 1)
        else:
            self.entries.append(("conv2d_out/kernel", (1, 1, c9, self.output_channels), "conv_kernel"))
            self.entries.append(("conv2d_out/bias", (self.output_channels,), "conv_bias"))

    def names(self):
        return [e[0] for e in self.entries]

    def init_params(self, seed: int = 0, dtype=torch.float32):
        """Glorot-uniform kernels, zero biases, BN gamma=1 beta=0 mean=0 var=1
        (Appendix B items 4, 10).  The reference leaves the UNet unseeded, so the
        actual stream is ours: one torch.Generator(seed), creation order."""
        gen = torch.Generator().manual_seed(seed)
        p = OrderedDict()
        for name, shape, kind in self.entries:
            if kind == "conv_kernel":
                kh, kw, ci, co = shape
                t = L.glorot_uniform(shape, gen, kh * kw * ci, kh * kw * co)
            elif kind == "convT_kernel":
                kh, kw, co, ci = shape
                # keras compute_fans on (kh,kw,Cout,Cin): fan_in = kh*kw*Cout, fan_out = kh*kw*Cin
                t = L.glorot_uniform(shape, gen, kh * kw * co, kh * kw * ci)
            elif kind in ("bn_gamma", "bn_var"):
                t = torch.ones(shape)
            else:
                t = torch.zeros(shape)
            p[name] = t.to(dtype)
        return p

    def trainable_names(self):
        return [n for n, _, k in self.entries if k not in ("bn_mean", "bn_var")]


class _Ctx:
    """Cursor over the parameter dict in creation order during one forward."""

    def __init__(self, params, training, new_stats):
        self.p = params
        self.training = training
        self.new_stats = new_stats
        self.nconv = 0
        self.nbn = 0
        self.nct = 0

    def conv(self, x, k):
        self.nconv += 1
        return L.conv2d(x, self.p[f"conv2d_{self.nconv}/kernel"], None, 1, "same")

    def bn(self, x, scale):
        self.nbn += 1
        n = f"batch_normalization_{self.nbn}"
        gamma = self.p[n + "/gamma"] if scale else None
        y, m, v = L.batch_norm(x, gamma, self.p[n + "/beta"], self.p[n + "/moving_mean"],
                               self.p[n + "/moving_variance"], self.training)
        if self.training:
            self.new_stats[n + "/moving_mean"] = m
            self.new_stats[n + "/moving_variance"] = v
        return y

    def convT(self, x):
        self.nct += 1
        n = f"conv2d_transpose_{self.nct}"
        return L.st(L.conv2d_transpose(x, self.p[n + "/kernel"], self.p[n + "/bias"], 2))


def _conv_bn(ctx, x, k, act, store_act=True):
    # L.st(...) marks the tensors the CUDA path writes to memory (identity unless L.storage(dtype) is active)
    x = L.st(ctx.conv(x, k))
    x = ctx.bn(x, False)
    if act == "relu":
        x = torch.relu(x)
        if store_act:
            x = L.st(x)
    elif act == "sigmoid":
        x = torch.sigmoid(x)
    return x


def _mres(ctx, inp):
    shortcut = _conv_bn(ctx, inp, 1, None)
    a = _conv_bn(ctx, inp, 3, "relu")
    b = _conv_bn(ctx, a, 3, "relu")
    c = _conv_bn(ctx, b, 3, "relu")
    out = torch.cat([a, b, c], dim=3)
    out = ctx.bn(out, True)
    out = L.st(torch.relu(shortcut + out))
    out = L.st(ctx.bn(out, True))
    return out


def _respath(ctx, length, x):
    for _ in range(length):
        shortcut = _conv_bn(ctx, x, 1, None)
        out = _conv_bn(ctx, x, 3, "relu", store_act=False)
        out = L.st(torch.relu(shortcut + out))
        x = L.st(ctx.bn(out, True))
    return x


def unet_forward(x, params, training: bool = False, output_channels: int = 1, taps=None):
    """model(x, training=...) for multi_res_unet.  x NHWC float.  Returns
    (y, new_moving_stats dict).  `taps` (optional dict) receives named
    intermediate activations for stage-level parity tests."""
    new_stats = {}
    ctx = _Ctx(params, training, new_stats)
    H, W = x.shape[1], x.shape[2]
    ph = (16 - H % 16) % 16
    pw = (16 - W % 16) % 16
    xp = L.reflection_pad(x, pw, ph)

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    m1 = tap("mres1", _mres(ctx, xp))
    p1 = L.st(L.max_pool_2x2(m1))
    r1 = tap("rp1", _respath(ctx, 4, m1))
    m2 = tap("mres2", _mres(ctx, p1))
    p2 = L.st(L.max_pool_2x2(m2))
    r2 = tap("rp2", _respath(ctx, 3, m2))
    m3 = tap("mres3", _mres(ctx, p2))
    p3 = L.st(L.max_pool_2x2(m3))
    r3 = tap("rp3", _respath(ctx, 2, m3))
    m4 = tap("mres4", _mres(ctx, p3))
    p4 = L.st(L.max_pool_2x2(m4))
    r4 = tap("rp4", _respath(ctx, 1, m4))
    m5 = tap("mres5", _mres(ctx, p4))
    u6 = torch.cat([ctx.convT(m5), r4], dim=3)
    m6 = tap("mres6", _mres(ctx, u6))
    u7 = torch.cat([ctx.convT(m6), r3], dim=3)
    m7 = tap("mres7", _mres(ctx, u7))
    u8 = torch.cat([ctx.convT(m7), r2], dim=3)
    m8 = tap("mres8", _mres(ctx, u8))
    u9 = torch.cat([ctx.convT(m8), r1], dim=3)
    m9 = tap("mres9", _mres(ctx, u9))
    t, b = ph // 2, ph // 2 + ph % 2
    l, r = pw // 2, pw // 2 + pw % 2
    cropped = m9[:, t:m9.shape[1] - b, l:m9.shape[2] - r, :]
    if output_channels == 1:
        y = _conv_bn(ctx, cropped, 1, "sigmoid")
    else:
        y = L.conv2d(cropped, params["conv2d_out/kernel"], params["conv2d_out/bias"], 1, "same")
        y = torch.softmax(y, dim=-1)
    return y, new_stats


class UNetTrainer:
    """One Keras TorchTrainer.train_step at a time."""

    def __init__(self, spec: UNetSpec, params, weighting: float, lr: float = 1e-3):
        self.spec = spec
        self.params = OrderedDict((k, v.clone()) for k, v in params.items())
        self.weighting = float(weighting)
        self.train_names = spec.trainable_names()
        for n in self.train_names:
            self.params[n].requires_grad_(True)
        self.opt = L.KerasAdam([self.params[n] for n in self.train_names], lr=lr)
        self.last_grads = None

    def train_step(self, x, y_true):
        for n in self.train_names:
            self.params[n].grad = None
        y_pred, new_stats = unet_forward(x, self.params, training=True)
        loss = L.weighted_bce(y_true, y_pred, self.weighting)
        loss.backward()
        grads = [self.params[n].grad for n in self.train_names]
        self.last_grads = OrderedDict((n, g.clone()) for n, g in zip(self.train_names, grads))
        self.opt.apply(grads, [self.params[n] for n in self.train_names])
        with torch.no_grad():
            for k, v in new_stats.items():
                self.params[k].copy_(v)
            yp = y_pred.detach()
            mae = (y_true - yp).abs().mean().item()
            acc = ((yp > 0.5).to(y_true.dtype) == y_true).to(torch.float32).mean().item()
        return {"loss": float(loss.item()), "mae": mae, "acc": acc}, y_pred.detach()


def synthetic_batch(n: int, h: int = 256, w: int = 256):
    """BASELINE config 2 inputs: x ~ U[0,1) seed 0,
    y = (U[0,1) seed 1 < 0.10), weighting = #zeros/#ones."""
    gx = torch.Generator().manual_seed(0)
    gy = torch.Generator().manual_seed(1)
    x = torch.rand(n, h, w, 1, generator=gx)
    y = (torch.rand(n, h, w, 1, generator=gy) < 0.10).to(torch.float32)
    ones = float(y.sum().item())
    weighting = float(y.numel() - ones) / max(ones, 1.0)
    return x, y, weighting
