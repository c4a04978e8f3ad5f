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
from .engine import AffineOp, Buf, ConvOp, Engine, Layout, NoiseOp, NormOp, PadCropOp, ParamSpec, UpsampleOp, View, pad8
from .nets import IN_EPS


def same_pad_lead(size: int, k: int, stride: int) -> int:
    """leading pad of Keras/TF 'same' for stride>1"""
    total = max((k - 1) - (size - 1) % stride, 0)
    return total // 2


class GT:
    """Symbolic activation of a GAN graph: view + spatial size + logical channel count (the reference's Keras tensor)."""

    def __init__(self, view: View, h: int, w: int, c: int):
        self.view, self.h, self.w, self.c = view, h, w, c

    @property
    def shape(self):
        return (None, self.h, self.w, self.c)


class _Base:
    def __init__(self, eng: Engine, prefix: str):
        self.e, self.prefix = eng, prefix
        self.creation_names: List[str] = []
        eng._builder = self
        self._nres = self._ndown = self._nup = 0

    def conv_norm_act(self, x: GT, wname: str, cout_l: int, k: int, stride: int, pad, pad_mode: int, nname: str, act: int,
                      transposed: bool = False, residual: Optional[View] = None, hw_out=None) -> GT:
        """Conv2D / Conv2DTranspose (no bias) -> GroupNormalization(groups=-1) -> activation [+ residual] (CycleGAN.py:323-358)."""
        e = self.e
        hw_out = hw_out or (x.h, x.w)
        norm = self.inorm(nname, cout_l, hw_out[0] * hw_out[1])
        raw = e.new_buf(hw_out[0], hw_out[1], pad8(cout_l), f"{self.prefix}{nname}_raw")
        e.add_op(ConvOp(e, x.view, raw.view(), (x.h, x.w), hw_out, wname, None, k, stride, pad, pad_mode, transposed,
                        stats=norm.stats_ref()))
        e.add_op(norm)
        out = e.new_buf(hw_out[0], hw_out[1], pad8(cout_l), f"{self.prefix}{nname}_out")
        e.add_op(AffineOp(e, hw_out[0] * hw_out[1], raw.view(), norm, residual, None, out.view(), act))
        return GT(out.view(), hw_out[0], hw_out[1], cout_l)

    # ---- the reference's block functions (CycleGAN.py:323-358) on symbolic tensors -------------------------------------
    def residual_block(self, x: GT, act: int = L.ACT_RELU) -> GT:
        i, f = self._nres, x.c
        self._nres += 1
        y = self.conv_norm_act(x, self.conv_w(f"res{i}_0", 3, f, f), f, 3, 1, (1, 1), L.PAD_REFLECT, f"res{i}_0_in", act)
        return self.conv_norm_act(y, self.conv_w(f"res{i}_1", 3, f, f), f, 3, 1, (1, 1), L.PAD_REFLECT, f"res{i}_1_in", L.ACT_NONE,
                                  residual=x.view)

    def downsample(self, x: GT, filters: int, act: int = L.ACT_RELU, k: int = 3, padding: str = "same", name: Optional[str] = None) -> GT:
        name = name or f"down{self._ndown}"
        self._ndown += 1
        if padding == "same":
            oh, ow = -(-x.h // 2), -(-x.w // 2)
            pad = (same_pad_lead(x.h, k, 2), same_pad_lead(x.w, k, 2))
        else:
            oh, ow = (x.h - k) // 2 + 1, (x.w - k) // 2 + 1
            pad = (0, 0)
        return self.conv_norm_act(x, self.conv_w(name, k, x.c, filters), filters, k, 2, pad, L.PAD_ZERO, f"{name}_in", act, hw_out=(oh, ow))

    def upsample(self, x: GT, filters: int, act: int = L.ACT_RELU, use_resize_convolution: bool = False) -> GT:
        i, e, f = self._nup, self.e, x.c
        self._nup += 1
        H, W = x.h, x.w
        if use_resize_convolution:
            # UpSampling2D(nearest) -> ReflectionPadding2D(1) -> Conv2D(3x3, valid, no bias) (CycleGAN.py:348-351)
            up = e.new_buf(2 * H, 2 * W, pad8(f), f"{self.prefix}up{i}_nearest")
            e.add_op(UpsampleOp(e, x.view, up.view(), H, W))
            return self.conv_norm_act(GT(up.view(), 2 * H, 2 * W, f), self.conv_w(f"up{i}", 3, f, filters), filters, 3, 1, (1, 1),
                                      L.PAD_REFLECT, f"up{i}_in", act)
        # Conv2DTranspose(3x3, s2, 'same') == torch padding 1, output_padding 1: the equivalent strided conv has pad 1
        return self.conv_norm_act(x, self.convT_w(f"up{i}", 3, f, filters), filters, 3, 2, (1, 1), L.PAD_ZERO, f"up{i}_in", act,
                                  transposed=True, hw_out=(2 * H, 2 * W))

    def _p(self, name, kind, lshape, pshape, maps, init="zeros", fans=(1, 1)):
        full = f"{self.prefix}{name}"
        self.e.add_param(ParamSpec(full, kind, lshape, pshape, maps, True, init, fans))
        self.creation_names.append(full)
        return full

    def conv_w(self, name, k, cin_l, cout_l):
        ci, co = pad8(cin_l), pad8(cout_l)
        return self._p(name + "/kernel", "conv_kernel", (k, k, cin_l, cout_l), (k, k, ci, co),
                       {2: np.arange(cin_l), 3: np.arange(cout_l)}, "glorot", (k * k * cin_l, k * k * cout_l))

    def conv_w_layout(self, name, k, lin: Layout, lout: Layout):
        """Conv2D kernel whose input / output channels live in segmented (concat) physical layouts."""
        return self._p(name + "/kernel", "conv_kernel", (k, k, lin.logical, lout.logical), (k, k, lin.phys, lout.phys),
                       {2: lin.index_map(), 3: lout.index_map()}, "glorot", (k * k * lin.logical, k * k * lout.logical))

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
                 prefix: str = "", in_buf: Optional[Buf] = None, use_skip_connection: bool = False,
                 use_resize_convolution: bool = False, **_ignored):
        super().__init__(eng, prefix)
        e = eng
        m = 2 ** n_down
        if use_skip_connection and (h % m or w % m):
            # the reference adds the (un-cropped, CycleGAN.py:394) padded output to a branch of the un-padded input
            raise ValueError(f"use_skip_connection needs an input size divisible by {m} (got {h}x{w}); Keras fails likewise")
        ph, pw = (m - h % m) % m, (m - w % m) % m
        self.in_buf = in_buf if in_buf is not None else e.new_buf(h, w, 8, prefix + "input", requires_grad=False)
        x, H, W = self.in_buf.view(), h, w
        if ph or pw:
            padded = e.new_buf(h + ph, w + pw, 8, prefix + "input_fix", requires_grad=self.in_buf.requires_grad)
            e.add_op(PadCropOp(e, x, padded.view(), (h, w), (h + ph, w + pw), ph // 2, pw // 2, "reflect"))
            x, H, W = padded.view(), h + ph, w + pw
        f = filters
        t = GT(x, H, W, 1)
        # stem: ReflectionPadding2D(3) + 7x7 valid, no bias, IN, relu
        t = self.conv_norm_act(t, self.conv_w("stem", 7, 1, f), f, 7, 1, (3, 3), L.PAD_REFLECT, "stem_in", L.ACT_RELU)
        for _ in range(n_down):
            f *= 2
            t = self.downsample(t, f)
        for _ in range(n_res):
            t = self.residual_block(t)
        for _ in range(n_up):
            f //= 2
            t = self.upsample(t, f, use_resize_convolution=use_resize_convolution)
        x, H, W = t.view, t.h, t.w
        wh = self.conv_w("head", 7, f, 1)
        bh = self.vec("head/bias", 1, "zeros")
        if use_skip_connection:
            # CycleGAN.py:396-415: [relu(IN(conv1x1(img))) + relu(IN(conv3x3(reflect_pad(img))))] -> IN -> relu, concatenated
            # with the head conv's output, then a bias-free 1x1 conv back to one channel.
            img = self.in_buf.view()
            fp = pad8(f)
            cat = e.new_buf(H, W, fp + 8, prefix + "skip_cat")
            s_act = self.conv_norm_act(GT(img, H, W, 1), self.conv_w("skip_short", 1, 1, f), f, 1, 1, (0, 0), L.PAD_ZERO, "skip_short_in",
                                       L.ACT_RELU).view
            w_o = self.conv_w("skip_conv", 3, 1, f)
            n_o = self.inorm("skip_conv_in", f, H * W)
            o_raw = e.new_buf(H, W, fp, prefix + "skip_conv_raw")
            e.add_op(ConvOp(e, img, o_raw.view(), (H, W), (H, W), w_o, None, 3, 1, (1, 1), L.PAD_REFLECT, False, stats=n_o.stats_ref()))
            e.add_op(n_o)
            n_s = self.inorm("skip_sum_in", f, H * W)
            summed = e.new_buf(H, W, fp, prefix + "skip_sum")
            e.add_op(AffineOp(e, H * W, s_act, None, o_raw.view(), n_o, summed.view(), L.ACT_NONE, actb=L.ACT_RELU,
                              stats_out=n_s.stats_ref()))
            e.add_op(n_s)
            e.add_op(AffineOp(e, H * W, summed.view(), n_s, None, None, cat.view(0, fp), L.ACT_RELU))
            e.add_op(ConvOp(e, x, cat.view(fp, 8), (H, W), (H, W), wh, bh, 7, 1, (3, 3), L.PAD_REFLECT, False))
            lcat = Layout([(f, fp), (1, 8)])
            wf = self.conv_w_layout("skip_out", 1, lcat, Layout.simple(1))
            raw = e.new_buf(H, W, 8, prefix + "head_raw")
            e.add_op(ConvOp(e, cat.view(), raw.view(), (H, W), (H, W), wf, None, 1, 1, (0, 0), L.PAD_ZERO, False))
        else:
            raw = e.new_buf(H, W, 8, prefix + "head_raw")
            e.add_op(ConvOp(e, x, raw.view(), (H, W), (H, W), wh, bh, 7, 1, (3, 3), L.PAD_REFLECT, False))
        self.out_buf = e.new_buf(H, W, 8, prefix + "output")
        e.add_op(AffineOp(e, H * W, raw.view(), None, None, None, self.out_buf.view(), L.ACT_TANH))
        self.out_hw = (H, W)


class DiscriminatorBuilder(_Base):
    def __init__(self, eng: Engine, h: int, w: int, filters: int = 128, n_down: int = 2, prefix: str = "",
                 in_buf: Optional[Buf] = None, gaussian_noise: float = 0.0, **_ignored):
        super().__init__(eng, prefix)
        e = eng
        if n_down > 3:
            raise NotImplementedError("PatchGAN blocks past the third use stride 1 (CycleGAN.py:441-444): not built")
        self.in_buf = in_buf if in_buf is not None else e.new_buf(h, w, 8, prefix + "input", requires_grad=False)
        x, H, W = self.in_buf.view(), h, w
        f = filters
        self.noise_ops: List[NoiseOp] = []

        def noisy(x, H, W, tag):
            """GaussianNoise(gaussian_noise_value) in front of every conv of the discriminator (CycleGAN.py:427-447)."""
            if gaussian_noise <= 0:
                return x
            out = e.new_buf(H, W, x.C, f"{prefix}{tag}_noisy", requires_grad=x.requires_grad)
            op = NoiseOp(e, x, out.view(), H * W, gaussian_noise, c_logical=(1 if tag == "d0" else x.C))
            self.noise_ops.append(e.add_op(op))
            return out.view()

        x = noisy(x, H, W, "d0")
        oh, ow = (H - 4) // 2 + 1, (W - 4) // 2 + 1
        w0, b0 = self.conv_w("d0", 4, 1, f), self.vec("d0/bias", f, "zeros")
        raw = e.new_buf(oh, ow, pad8(f), prefix + "d0_raw")
        e.add_op(ConvOp(e, x, raw.view(), (H, W), (oh, ow), w0, b0, 4, 2, (0, 0), L.PAD_ZERO, False))
        act = e.new_buf(oh, ow, pad8(f), prefix + "d0_out")
        e.add_op(AffineOp(e, oh * ow, raw.view(), None, None, None, act.view(), L.ACT_LEAKY))
        x, H, W = act.view(), oh, ow
        for i in range(n_down):
            oh, ow = (H - 4) // 2 + 1, (W - 4) // 2 + 1
            x = noisy(x, H, W, f"d{i + 1}")
            wn = self.conv_w(f"d{i + 1}", 4, f, 2 * f)
            norm = self.inorm(f"d{i + 1}_in", 2 * f, oh * ow)
            raw = e.new_buf(oh, ow, pad8(2 * f), f"{prefix}d{i + 1}_raw")
            e.add_op(ConvOp(e, x, raw.view(), (H, W), (oh, ow), wn, None, 4, 2, (0, 0), L.PAD_ZERO, False, stats=norm.stats_ref()))
            e.add_op(norm)
            act = e.new_buf(oh, ow, pad8(2 * f), f"{prefix}d{i + 1}_out")
            e.add_op(AffineOp(e, oh * ow, raw.view(), norm, None, None, act.view(), L.ACT_LEAKY))
            x, H, W, f = act.view(), oh, ow, 2 * f
        x = noisy(x, H, W, "out")
        oh, ow = H - 3, W - 3
        wo, bo = self.conv_w("out", 4, f, 1), self.vec("out/bias", 1, "zeros")
        self.out_buf = e.new_buf(oh, ow, 8, prefix + "output")
        e.add_op(ConvOp(e, x, self.out_buf.view(), (H, W), (oh, ow), wo, bo, 4, 1, (0, 0), L.PAD_ZERO, False))
        self.out_hw = (oh, ow)
