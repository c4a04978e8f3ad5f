"""CycleGAN ResNet generator and PatchGAN discriminator recorded as engine ops.

/root/reference/Releases/Version 1.2.0/CycleGAN.py: residual_block :323-337, downsample :339-345, upsample :347-358,
get_resnet_generator :360-423, get_discriminator :425-451.  Configuration = the one StartProcess.py drives
(use_skip_connection=False, gaussian_noise_value=0, use_resize_convolution=False, tanh output).

InstanceNorm (GroupNormalization(groups=-1)) moments are accumulated per (sample, channel) in the conv epilogues;
the 512->512 3x3 residual convs (86 % of the FLOPs) run on tcgen05 with the reflection padding folded into the
halo-tile loader; strided / 7x7 / 4x4 / transposed layers use the CUDA-core kernels.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import _lib as L
from .engine import AffineOp, Buf, ConvOp, Engine, Layout, NormOp, PadCropOp, ParamSpec, View, pad8
from .nets import IN_EPS


def same_pad_lead(size: int, k: int, stride: int) -> int:
    """leading pad of Keras/TF 'same' for stride>1"""
    total = max((k - 1) - (size - 1) % stride, 0)
    return total // 2


class _Base:
    def __init__(self, eng: Engine, prefix: str):
        self.e, self.prefix = eng, prefix
        self.creation_names: List[str] = []

    def _p(self, name, kind, lshape, pshape, maps, init="zeros", fans=(1, 1)):
        full = f"{self.prefix}{name}"
        self.e.add_param(ParamSpec(full, kind, lshape, pshape, maps, True, init, fans))
        self.creation_names.append(full)
        return full

    def conv_w(self, name, k, cin_l, cout_l):
        ci, co = pad8(cin_l), pad8(cout_l)
        return self._p(name + "/kernel", "conv_kernel", (k, k, cin_l, cout_l), (k, k, ci, co),
                       {2: np.arange(cin_l), 3: np.arange(cout_l)}, "glorot", (k * k * cin_l, k * k * cout_l))

    def convT_w(self, name, k, cin_l, cout_l):
        # Keras Conv2DTranspose kernel (kh,kw,Cout,Cin) == HWIO kernel of the equivalent strided conv (I=Cout_T, O=Cin_T)
        ci, co = pad8(cin_l), pad8(cout_l)
        return self._p(name + "/kernel", "convT_kernel", (k, k, cout_l, cin_l), (k, k, co, ci),
                       {2: np.arange(cout_l), 3: np.arange(cin_l)}, "glorot", (k * k * cout_l, k * k * cin_l))

    def vec(self, name, c_l, init):
        return self._p(name, "vector", (c_l,), (pad8(c_l),), {0: np.arange(c_l)}, init)

    def inorm(self, name, c_l, hw) -> NormOp:
        g = self.vec(name + "/gamma", c_l, "ones")
        b = self.vec(name + "/beta", c_l, "zeros")
        return NormOp(self.e, f"{self.prefix}{name}", pad8(c_l), hw, IN_EPS, g, b, None, 0.0, groups=self.e.N)


class GeneratorBuilder(_Base):
    def __init__(self, eng: Engine, h: int, w: int, filters: int = 64, n_down: int = 3, n_res: int = 9, n_up: int = 3,
                 prefix: str = "", in_buf: Optional[Buf] = None):
        super().__init__(eng, prefix)
        e = eng
        m = 2 ** n_down
        ph, pw = (m - h % m) % m, (m - w % m) % m
        self.in_buf = in_buf if in_buf is not None else e.new_buf(h, w, 8, prefix + "input", requires_grad=False)
        x, H, W = self.in_buf.view(), h, w
        if ph or pw:
            padded = e.new_buf(h + ph, w + pw, 8, prefix + "input_fix", requires_grad=self.in_buf.requires_grad)
            e.add_op(PadCropOp(e, x, padded.view(), (h, w), (h + ph, w + pw), ph // 2, pw // 2, "reflect"))
            x, H, W = padded.view(), h + ph, w + pw
        f = filters

        def conv_in_act(x, hw_in, hw_out, wname, k, stride, pad, pad_mode, cout_l, nname, act, transposed=False, residual=None):
            norm = self.inorm(nname, cout_l, hw_out[0] * hw_out[1])
            raw = e.new_buf(hw_out[0], hw_out[1], pad8(cout_l), f"{prefix}{nname}_raw")
            e.add_op(ConvOp(e, x, raw.view(), hw_in, hw_out, wname, None, k, stride, pad, pad_mode, transposed,
                            stats=norm.stats_ref()))
            e.add_op(norm)
            out = e.new_buf(hw_out[0], hw_out[1], pad8(cout_l), f"{prefix}{nname}_out")
            e.add_op(AffineOp(e, hw_out[0] * hw_out[1], raw.view(), norm, residual, None, out.view(), act))
            return out.view()

        # stem: ReflectionPadding2D(3) + 7x7 valid, no bias, IN, relu
        x = conv_in_act(x, (H, W), (H, W), self.conv_w("stem", 7, 1, f), 7, 1, (3, 3), L.PAD_REFLECT, f, "stem_in", L.ACT_RELU)
        for i in range(n_down):
            oh, ow = -(-H // 2), -(-W // 2)
            x = conv_in_act(x, (H, W), (oh, ow), self.conv_w(f"down{i}", 3, f, 2 * f), 3, 2,
                            (same_pad_lead(H, 3, 2), same_pad_lead(W, 3, 2)), L.PAD_ZERO, 2 * f, f"down{i}_in", L.ACT_RELU)
            f, H, W = 2 * f, oh, ow
        for i in range(n_res):
            y = conv_in_act(x, (H, W), (H, W), self.conv_w(f"res{i}_0", 3, f, f), 3, 1, (1, 1), L.PAD_REFLECT, f, f"res{i}_0_in", L.ACT_RELU)
            x = conv_in_act(y, (H, W), (H, W), self.conv_w(f"res{i}_1", 3, f, f), 3, 1, (1, 1), L.PAD_REFLECT, f, f"res{i}_1_in",
                            L.ACT_NONE, residual=x)
        for i in range(n_up):
            # Conv2DTranspose(3x3, s2, 'same') == torch padding 1, output_padding 1: the equivalent strided conv has pad 1
            x = conv_in_act(x, (H, W), (2 * H, 2 * W), self.convT_w(f"up{i}", 3, f, f // 2), 3, 2, (1, 1), L.PAD_ZERO, f // 2,
                            f"up{i}_in", L.ACT_RELU, transposed=True)
            f, H, W = f // 2, 2 * H, 2 * W
        wh = self.conv_w("head", 7, f, 1)
        bh = self.vec("head/bias", 1, "zeros")
        raw = e.new_buf(H, W, 8, prefix + "head_raw")
        e.add_op(ConvOp(e, x, raw.view(), (H, W), (H, W), wh, bh, 7, 1, (3, 3), L.PAD_REFLECT, False))
        self.out_buf = e.new_buf(H, W, 8, prefix + "output")
        e.add_op(AffineOp(e, H * W, raw.view(), None, None, None, self.out_buf.view(), L.ACT_TANH))
        self.out_hw = (H, W)


class DiscriminatorBuilder(_Base):
    def __init__(self, eng: Engine, h: int, w: int, filters: int = 128, n_down: int = 2, prefix: str = "",
                 in_buf: Optional[Buf] = None):
        super().__init__(eng, prefix)
        e = eng
        self.in_buf = in_buf if in_buf is not None else e.new_buf(h, w, 8, prefix + "input", requires_grad=False)
        x, H, W = self.in_buf.view(), h, w
        f = filters
        oh, ow = (H - 4) // 2 + 1, (W - 4) // 2 + 1
        w0, b0 = self.conv_w("d0", 4, 1, f), self.vec("d0/bias", f, "zeros")
        raw = e.new_buf(oh, ow, pad8(f), prefix + "d0_raw")
        e.add_op(ConvOp(e, x, raw.view(), (H, W), (oh, ow), w0, b0, 4, 2, (0, 0), L.PAD_ZERO, False))
        act = e.new_buf(oh, ow, pad8(f), prefix + "d0_out")
        e.add_op(AffineOp(e, oh * ow, raw.view(), None, None, None, act.view(), L.ACT_LEAKY))
        x, H, W = act.view(), oh, ow
        for i in range(n_down):
            oh, ow = (H - 4) // 2 + 1, (W - 4) // 2 + 1
            wn = self.conv_w(f"d{i + 1}", 4, f, 2 * f)
            norm = self.inorm(f"d{i + 1}_in", 2 * f, oh * ow)
            raw = e.new_buf(oh, ow, pad8(2 * f), f"{prefix}d{i + 1}_raw")
            e.add_op(ConvOp(e, x, raw.view(), (H, W), (oh, ow), wn, None, 4, 2, (0, 0), L.PAD_ZERO, False, stats=norm.stats_ref()))
            e.add_op(norm)
            act = e.new_buf(oh, ow, pad8(2 * f), f"{prefix}d{i + 1}_out")
            e.add_op(AffineOp(e, oh * ow, raw.view(), norm, None, None, act.view(), L.ACT_LEAKY))
            x, H, W, f = act.view(), oh, ow, 2 * f
        oh, ow = H - 3, W - 3
        wo, bo = self.conv_w("out", 4, f, 1), self.vec("out/bias", 1, "zeros")
        self.out_buf = e.new_buf(oh, ow, 8, prefix + "output")
        e.add_op(ConvOp(e, x, self.out_buf.view(), (H, W), (oh, ow), wo, bo, 4, 1, (0, 0), L.PAD_ZERO, False))
        self.out_hw = (oh, ow)
