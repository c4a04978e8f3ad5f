"""WGAN-GP critic and generator recorded as engine ops.

/root/reference/Releases/Version 1.2.0/WassersteinGAN.py: conv_block :547-567, get_discriminator_model :569-621,
upsample_block :624-649, get_generator_model :651-684.

Critic: 4 x [Conv2D(5x5, stride 2, 'same', bias) -> LeakyReLU(0.2) -> (Dropout 0.3 after blocks 1 and 2)] -> Flatten ->
Dropout(0.2) -> Dense(1).  LeakyReLU and Dropout are one multiplicative mask (engine.MaskOp), which also gives the
LINEARISED critic the gradient penalty needs: a tower built with `like=<critic tower>` re-uses that tower's masks and runs
bias-free, so that back-propagating <grad_x D(x_hat), u> through it yields d(penalty)/d(weights) without a second-order graph.

Generator: Dense(n_z -> H/8*W/8*256, no bias) -> BatchNormalization -> LeakyReLU -> Reshape -> 3 x [UpSampling2D(2) ->
Conv2D(3x3, 'same', no bias) -> BatchNormalization -> LeakyReLU / tanh].

Dense layers are 1x1 convolutions over the batch laid out as ONE image row (n=1, h=1, w=batch): BatchNormalization of the
Dense output is then a plain per-channel norm over that row, and Flatten / Reshape are aliases of the same NHWC storage.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import _lib as L
from .engine import AffineOp, Buf, ConvOp, Engine, MaskOp, NormOp, ParamSpec, UpsampleOp, pad8
from .gan_nets import same_pad_lead
from .nets import BN_EPS, BN_MOMENTUM

CRITIC_FILTERS = (64, 128, 256, 512)
CRITIC_DROP = (0.0, 0.3, 0.3, 0.0)      # use_dropout of the four conv blocks (WassersteinGAN.py:571-614)
FLAT_DROP = 0.2                          # :618


class _Base:
    def __init__(self, eng: Engine, prefix: str):
        self.e, self.prefix = eng, prefix
        self.creation_names: List[str] = []
        eng._builder = self

    def _p(self, name, kind, lshape, pshape, maps, trainable=True, init="zeros", fans=(1, 1), to_phys_fn=None, to_logical_fn=None):
        full = f"{self.prefix}{name}"
        self.e.add_param(ParamSpec(full, kind, lshape, pshape, maps, trainable, init, fans, to_phys_fn=to_phys_fn, to_logical_fn=to_logical_fn))
        self.creation_names.append(full)
        return full

    def conv_w(self, name, k, cin_l, cout_l):
        return self._p(name + "/kernel", "conv_kernel", (k, k, cin_l, cout_l), (k, k, pad8(cin_l), pad8(cout_l)),
                       {2: np.arange(cin_l), 3: np.arange(cout_l)}, True, "glorot", (k * k * cin_l, k * k * cout_l))

    def dense_w(self, name, cin_l, cout_l):
        """keras Dense kernel (in, out) stored as the 1x1 conv kernel (1, 1, in8, out8)."""
        ci, co = pad8(cin_l), pad8(cout_l)

        def to_phys(a):
            out = np.zeros((1, 1, ci, co), dtype=np.float32)
            out[0, 0, :cin_l, :cout_l] = a
            return out

        return self._p(name + "/kernel", "dense_kernel", (cin_l, cout_l), (1, 1, ci, co), {}, True, "glorot", (cin_l, cout_l),
                       to_phys_fn=to_phys, to_logical_fn=lambda ph: ph[0, 0, :cin_l, :cout_l])

    def vec(self, name, c_l, init, trainable=True):
        return self._p(name, "vector", (c_l,), (pad8(c_l),), {0: np.arange(c_l)}, trainable, init)

    def bn(self, name, c_l, count) -> NormOp:
        g = self.vec(name + "/gamma", c_l, "ones")
        b = self.vec(name + "/beta", c_l, "zeros")
        mm = self.vec(name + "/moving_mean", c_l, "zeros", trainable=False)
        mv = self.vec(name + "/moving_variance", c_l, "ones", trainable=False)
        return NormOp(self.e, f"{self.prefix}{name}", pad8(c_l), count, BN_EPS, g, b, (mm, mv), BN_MOMENTUM)


class WganCriticBuilder(_Base):
    """get_discriminator_model (:569-621).  `like`: a builder of the same critic whose masks this tower re-uses, bias-free
    (the linearised critic of the gradient penalty; its input is the penalty's direction u, see wgan_model.py)."""

    def __init__(self, eng: Engine, h: int, w: int, prefix: str = "", in_buf: Optional[Buf] = None, like: Optional["WganCriticBuilder"] = None,
                 input_requires_grad: bool = False):
        super().__init__(eng, prefix)
        if h % 16 or w % 16:
            raise ValueError(f"the WGAN critic needs image sizes divisible by 16 (the reference pads to that, WassersteinGAN.py:343-352); got {h}x{w}")
        e, n = eng, eng.N
        self.in_buf = in_buf if in_buf is not None else e.new_buf(h, w, 8, prefix + "input", requires_grad=input_requires_grad)
        x, H, W, cin = self.in_buf.view(), h, w, 1
        self.masks: List[MaskOp] = []
        linear = like is not None
        for i, f in enumerate(CRITIC_FILTERS):
            oh, ow = -(-H // 2), -(-W // 2)
            wn = self.conv_w(f"c{i}", 5, cin, f)
            bn_ = self.vec(f"c{i}/bias", f, "zeros")
            raw = e.new_buf(oh, ow, pad8(f), f"{prefix}c{i}_raw")
            e.add_op(ConvOp(e, x, raw.view(), (H, W), (oh, ow), wn, None if linear else bn_, 5, 2,
                            (same_pad_lead(H, 5, 2), same_pad_lead(W, 5, 2)), L.PAD_ZERO, False))
            act = e.new_buf(oh, ow, pad8(f), f"{prefix}c{i}_out")
            if linear:
                op = MaskOp(e, raw.view(), act.view(), n * oh * ow, like=like.masks[i])
            else:
                op = MaskOp(e, raw.view(), act.view(), n * oh * ow, z=raw.view(), rate=CRITIC_DROP[i])
            self.masks.append(e.add_op(op))
            x, H, W, cin = act.view(), oh, ow, f
        # Flatten -> Dropout(0.2) -> Dense(1): the (n, H, W, 512) tensor seen as one image row of n pixels with H*W*512 channels
        feat = H * W * cin
        dropped = e.new_buf(H, W, cin, prefix + "flat_drop")
        if linear:
            op = MaskOp(e, x, dropped.view(), n * H * W, like=like.masks[4])
        else:
            op = MaskOp(e, x, dropped.view(), n * H * W, z=None, rate=FLAT_DROP)
        self.masks.append(e.add_op(op))
        flat = dropped.alias(1, 1, n, feat, prefix + "flat")
        wd = self.dense_w("dense", feat, 1)
        bd = self.vec("dense/bias", 1, "zeros")
        self.out_buf = e.new_buf(1, n, 8, prefix + "output", n=1)           # logits: one image row, pixel = sample, channel 0
        e.add_op(ConvOp(e, flat.view(), self.out_buf.view(), (1, n), (1, n), wd, None if linear else bd, 1, 1, (0, 0), L.PAD_ZERO, False, n=1))
        self.out_hw = (1, n)
        self.feat = feat


class WganGeneratorBuilder(_Base):
    """get_generator_model (:651-684)."""

    def __init__(self, eng: Engine, h: int, w: int, n_z: int = 128, prefix: str = ""):
        super().__init__(eng, prefix)
        if h % 8 or w % 8:
            raise ValueError(f"the WGAN generator needs image sizes divisible by 8; got {h}x{w}")
        e, n = eng, eng.N
        h8, w8 = h // 8, w // 8
        feat = h8 * w8 * 256
        self.n_z = n_z
        self.in_buf = e.new_buf(1, n, pad8(n_z), prefix + "latent", requires_grad=False, n=1)      # one image row, pixel = sample
        wd = self.dense_w("dense", n_z, feat)
        norm = self.bn("bn0", feat, n)
        raw = e.new_buf(1, n, feat, prefix + "dense_raw", n=1)
        e.add_op(ConvOp(e, self.in_buf.view(), raw.view(), (1, n), (1, n), wd, None, 1, 1, (0, 0), L.PAD_ZERO, False, stats=norm.stats_ref(), n=1))
        e.add_op(norm)
        act = e.new_buf(1, n, feat, prefix + "dense_out", n=1)
        for c0 in range(0, feat, 2048):          # the affine kernels stage at most 2048 channels of parameters per launch
            cc = min(2048, feat - c0)
            op = AffineOp(e, n, raw.view(c0, cc), norm, None, None, act.view(c0, cc), L.ACT_LEAKY, n=1)
            op.coff_a = c0
            e.add_op(op)
        x = act.alias(n, h8, w8, 256, prefix + "reshape").view()                                  # Reshape((H/8, W/8, 256))
        H, W, cin = h8, w8, 256
        for i, f in enumerate((128, 64, 1)):
            up = e.new_buf(2 * H, 2 * W, pad8(cin), f"{prefix}up{i}_nearest")
            e.add_op(UpsampleOp(e, x, up.view(), H, W))
            H, W = 2 * H, 2 * W
            wn = self.conv_w(f"up{i}", 3, cin, f)
            norm = self.bn(f"bn{i + 1}", f, n * H * W)
            raw = e.new_buf(H, W, pad8(f), f"{prefix}up{i}_raw")
            e.add_op(ConvOp(e, up.view(), raw.view(), (H, W), (H, W), wn, None, 3, 1, (1, 1), L.PAD_ZERO, False, stats=norm.stats_ref()))
            e.add_op(norm)
            out = e.new_buf(H, W, pad8(f), f"{prefix}up{i}_out")
            e.add_op(AffineOp(e, H * W, raw.view(), norm, None, None, out.view(), L.ACT_TANH if i == 2 else L.ACT_LEAKY))
            x, cin = out.view(), f
        self.out_buf = out
        self.out_hw = (H, W)
