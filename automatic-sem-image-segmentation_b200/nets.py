"""Network builders: record the reference's Keras graphs as engine ops.

MultiRes-UNet: /root/reference/Releases/Version 1.2.0/UNet_Segmentation.py:401-562
(conv2d_bn :401-426, multi_res_block :451-474, res_path :476-503, multi_res_unet :505-562).

Besides the ops, the builder records every Keras layer call in a small DAG (`KerasGraph`) so that
`get_weights()/set_weights()` can be ordered like `keras.Model.layers` (depth-sorted topological
order).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib as L
import os

from .engine import (AffineOp, Buf, ConvOp, Engine, Layout, NormOp, PadCropOp, PairConvOp, ParamSpec, PoolOp, SegView, ShuffleOp, View, pad8)

BN_MOMENTUM = 0.99
BN_EPS = 1e-3
IN_EPS = 1e-5


class KerasGraph:
    """The functional-API layer DAG, only to reproduce keras.Model.layers ordering [K3.5 function.map_graph]."""

    def __init__(self):
        self.inputs: List[List[int]] = []     # per layer: ids of the layers producing its inputs (call order)
        self.weights: List[List[str]] = []    # per layer: engine parameter names in Keras' layer.weights order
        self.names: List[str] = []

    def layer(self, name: str, inputs: List[int], weights: Optional[List[str]] = None) -> int:
        self.inputs.append(list(inputs))
        self.weights.append(list(weights or []))
        self.names.append(name)
        return len(self.names) - 1

    def layer_order(self, output: int) -> List[int]:
        # pre-order DFS from the output assigns the tie-break index; post-order gives decreasing depth
        op_index: Dict[int, int] = {}
        post: List[int] = []
        finished = set()
        stack = [(output, 0)]
        while stack:
            node, i = stack.pop()
            if i == 0:
                if node in finished:
                    continue
                if node not in op_index:
                    op_index[node] = len(op_index)
            ins = self.inputs[node]
            # skip already finished children
            while i < len(ins) and ins[i] in finished:
                i += 1
            if i < len(ins):
                stack.append((node, i + 1))
                stack.append((ins[i], 0))
            else:
                if node not in finished:
                    finished.add(node)
                    post.append(node)
        depth = {n: 0 for n in post}
        for node in reversed(post):
            for parent in self.inputs[node]:
                depth[parent] = max(depth[parent], depth[node] + 1)
        return sorted(post, key=lambda n: (-depth[n], op_index[n]))

    def weight_names(self, output: int) -> List[str]:
        out = []
        for l in self.layer_order(output):
            out.extend(self.weights[l])
        return out


class T:
    """A symbolic activation: a View plus spatial size, channel layout and the Keras layer that produced it."""

    def __init__(self, view: View, h: int, w: int, layout: Layout, klayer: int):
        self.view, self.h, self.w, self.layout, self.klayer = view, h, w, layout, klayer


class UNetBuilder:
    def __init__(self, eng: Engine, h: int, w: int, filters: int = 16, in_channels: int = 1, output_channels: int = 1,
                 build: bool = True):
        if output_channels != 1:
            raise NotImplementedError("softmax head (output_channels>1) is not on the north-star path yet")
        self.e, self.h, self.w, self.filters = eng, h, w, filters
        self.kg = KerasGraph()
        self.nconv = self.nbn = self.nct = 0
        self.creation_names: List[str] = []
        self.taps: Dict[str, T] = {}
        eng._builder = self          # the reference's static layer functions find their graph through the tensor's engine
        self._anon = 0
        if build:
            self._build(in_channels)
        else:
            # functional use (UNet.conv2d_bn / multi_res_block / res_path called on keras_compat.Input): only the input exists
            lin = Layout.simple(in_channels)
            self.in_buf = eng.new_buf(h, w, lin.phys, "input", requires_grad=False)
            self.input = T(self.in_buf.view(), h, w, lin, self.kg.layer("input", []))
            self.out_buf = self.output = self.keras_output_layer = None

    # ---- generic layer functions behind the reference's statics (UNet_Segmentation.py:401-449) -------------------------
    def conv2d_bn(self, x: T, filters: int, k: int, activation: Optional[str]) -> T:
        """Conv2D(k x k, same, no bias) -> BatchNormalization(scale=False) -> optional activation, materialised."""
        act = {None: L.ACT_NONE, "relu": L.ACT_RELU, "sigmoid": L.ACT_SIGMOID}[activation]
        lay = Layout.simple(filters)
        raw, bn = self.conv2d_bn_raw(x, lay, k)
        klast = raw.klayer
        if activation is not None:
            klast = self.kg.layer("activation", [raw.klayer])
        out = self.e.new_buf(x.h, x.w, lay.phys, f"conv{self.nconv}_out")
        self.e.add_op(AffineOp(self.e, x.h * x.w, raw.view, bn, None, None, out.view(), act))
        return T(out.view(), x.h, x.w, lay, klast)

    def anon(self, prefix: str) -> str:
        self._anon += 1
        return f"{prefix}_{self._anon}"

    def set_output(self, t: T):
        self.output, self.out_buf, self.keras_output_layer = t, t.view.buf, t.klayer

    # ---- parameter helpers --------------------------------------------------------------------------
    def _add(self, spec: ParamSpec):
        self.e.add_param(spec)
        self.creation_names.append(spec.name)

    def _conv_param(self, lin: Layout, lout: Layout, k: int) -> str:
        self.nconv += 1
        name = f"conv2d_{self.nconv}/kernel"
        self._add(ParamSpec(name, "conv_kernel", (k, k, lin.logical, lout.logical), (k, k, lin.phys, lout.phys),
                            {2: lin.index_map(), 3: lout.index_map()}, True, "glorot",
                            (k * k * lin.logical, k * k * lout.logical)))
        return name

    def _bn_params(self, lay: Layout, scale: bool):
        self.nbn += 1
        base = f"batch_normalization_{self.nbn}"
        m = {0: lay.index_map()}
        names = {}
        if scale:
            names["gamma"] = base + "/gamma"
            self._add(ParamSpec(names["gamma"], "bn_gamma", (lay.logical,), (lay.phys,), m, True, "ones"))
        names["beta"] = base + "/beta"
        self._add(ParamSpec(names["beta"], "bn_beta", (lay.logical,), (lay.phys,), m, True, "zeros"))
        names["mean"] = base + "/moving_mean"
        self._add(ParamSpec(names["mean"], "bn_mean", (lay.logical,), (lay.phys,), m, False, "zeros"))
        names["var"] = base + "/moving_variance"
        self._add(ParamSpec(names["var"], "bn_var", (lay.logical,), (lay.phys,), m, False, "ones"))
        return base, names

    def _bn(self, lay: Layout, scale: bool, count: int) -> Tuple[NormOp, List[str]]:
        base, n = self._bn_params(lay, scale)
        norm = NormOp(self.e, base, lay.phys, count, BN_EPS, n.get("gamma"), n["beta"], (n["mean"], n["var"]), BN_MOMENTUM)
        wl = ([n["gamma"]] if scale else []) + [n["beta"], n["mean"], n["var"]]
        return norm, wl

    # ---- layer helpers (names follow the reference's functions) -------------------------------------------
    def conv2d_bn_raw(self, x: T, lout: Layout, k: int, out_view: Optional[View] = None):
        """Conv2D(no bias) + moments of BatchNormalization(scale=False); returns raw conv output + its NormOp."""
        e = self.e
        w = self._conv_param(x.layout, lout, k)
        kconv = self.kg.layer(f"conv2d_{self.nconv}", [x.klayer], [w])
        if out_view is None:
            out_view = e.new_buf(x.h, x.w, lout.phys, f"conv{self.nconv}_raw").view()
        norm, wl = self._bn(lout, False, e.N * x.h * x.w)
        kbn = self.kg.layer(norm.name, [kconv], wl)
        e.add_op(ConvOp(e, x.view, out_view, (x.h, x.w), (x.h, x.w), w, None, k, 1, (k // 2, k // 2), L.PAD_ZERO, False,
                        stats=norm.stats_ref()))
        e.add_op(norm)
        return T(out_view, x.h, x.w, lout, kbn), norm

    def res_unit_raw(self, x: T, lay: Layout):
        """The two convs of one res_path unit (UNet_Segmentation.py:490-499: 1x1 shortcut, then 3x3; both conv2d_bn without
        activation at this point) -> (shortcut raw, its norm, 3x3 raw, its norm).  In bf16 tensor-core mode, where the merged
        weight image stays resident in shared memory, they run as ONE 3x3 conv with 2C outputs (engine.PairConvOp): one
        launch and one read of x instead of two, one data gradient without a read-modify-write of dx; at N <= 64 the wider
        MMA costs the tensor pipe nothing.  Names / creation order of the four Keras layers are unchanged."""
        e = self.e
        c = lay.phys
        cin16 = (x.layout.phys + 15) // 16 * 16
        if not (e.tc_enabled and 9 * cin16 * 2 * c * 2 <= 48 * 1024 and os.environ.get("SEMB_NO_PAIR_CONV") is None):
            s_raw, bns = self.conv2d_bn_raw(x, lay, 1)
            o_raw, bno = self.conv2d_bn_raw(x, lay, 3)
            return s_raw, bns, o_raw, bno
        count = e.N * x.h * x.w
        ws = self._conv_param(x.layout, lay, 1)
        ks = self.kg.layer(f"conv2d_{self.nconv}", [x.klayer], [ws])
        bns, wls = self._bn(lay, False, count)
        kbns = self.kg.layer(bns.name, [ks], wls)
        wo = self._conv_param(x.layout, lay, 3)
        ko = self.kg.layer(f"conv2d_{self.nconv}", [x.klayer], [wo])
        bno, wlo = self._bn(lay, False, count)
        kbno = self.kg.layer(bno.name, [ko], wlo)
        buf = e.new_buf(x.h, x.w, 2 * c, f"conv{self.nconv}_pair_raw")            # [3x3 output | shortcut output]
        mom = e.zeroed.add(f"{bno.name}/pair_moments", 4 * 2 * c)                  # fp64 [sum 2c | sum of squares 2c]
        bno.bind_stats(mom, 0, 2 * c)
        bns.bind_stats(mom, c, 2 * c)
        e.add_op(PairConvOp(e, x.view, buf.view(), (x.h, x.w), wo, ws, c, c, stats=(mom, 0, 0, 2 * c)))
        e.add_op(bns)
        e.add_op(bno)
        return T(buf.view(c, c), x.h, x.w, lay, kbns), bns, T(buf.view(0, c), x.h, x.w, lay, kbno), bno

    def multi_res_block(self, u: int, inp: T, name: Optional[str] = None, alpha: float = 1.67) -> T:
        """UNet_Segmentation.py:451-474."""
        e = self.e
        name = name or self.anon("mres")
        wdt = alpha * u
        ca, cb, cc = int(wdt * 0.167), int(wdt * 0.333), int(wdt * 0.5)
        la, lb, lc = Layout.simple(ca), Layout.simple(cb), Layout.simple(cc)
        lcat = Layout.concat(la, lb, lc)
        hw = inp.h * inp.w
        count = e.N * hw

        s_raw, bn0 = self.conv2d_bn_raw(inp, lcat, 1)                       # shortcut (activation=None)
        # The three branches (raw conv outputs AND their activations) live in their OWN compact buffers: a 16-byte slice of a
        # 64-byte pixel costs the full DRAM sector (round 1: 4x read amplification of the convs that consumed 8-channel
        # slices of the concat buffer).  `concatenate([a, b, c])` never materialises: the block's add + ReLU kernel reads the
        # three tensors as segments of its second operand (semb_affine_desc.nseg_b).
        offs = [0, la.phys, la.phys + lb.phys]
        # The BatchNormalization over the concat is created after the three conv2d_bn in the reference (creation
        # order matters for variable names); its moments are accumulated by the three activation passes, so their
        # stats_out reference is bound once that NormOp exists.
        x = inp
        acts, act_ops = [], []
        for lay, off in zip((la, lb, lc), offs):
            raw, bn = self.conv2d_bn_raw(x, lay, 3)
            kact = self.kg.layer("activation", [raw.klayer])
            act_view = e.new_buf(inp.h, inp.w, lay.phys, f"{name}_act{len(acts)}").view()
            act_ops.append((e.add_op(AffineOp(e, hw, raw.view, bn, None, None, act_view, L.ACT_RELU)), off))
            x = T(act_view, inp.h, inp.w, lay, kact)
            acts.append(x)
        kcat = self.kg.layer("concatenate", [a.klayer for a in acts])
        bn4, wl4 = self._bn(lcat, True, count)
        kbn4 = self.kg.layer(bn4.name, [kcat], wl4)
        for op, off in act_ops:
            op.stats_out = bn4.stats_ref(off)
        e.add_op(bn4)
        kadd = self.kg.layer("add", [s_raw.klayer, kbn4])
        kact = self.kg.layer("activation", [kadd])
        bn5, wl5 = self._bn(lcat, True, count)
        kbn5 = self.kg.layer(bn5.name, [kact], wl5)
        out1 = e.new_buf(inp.h, inp.w, lcat.phys, name + "_sum")
        e.add_op(AffineOp(e, hw, s_raw.view, bn0, SegView([a.view for a in acts]), bn4, out1.view(), L.ACT_RELU, stats_out=bn5.stats_ref()))
        e.add_op(bn5)
        out = e.new_buf(inp.h, inp.w, lcat.phys, name)
        e.add_op(AffineOp(e, hw, out1.view(), bn5, None, None, out.view(), L.ACT_NONE))
        t = T(out.view(), inp.h, inp.w, lcat, kbn5)
        self.taps[name] = t
        return t

    def res_path(self, filters: int, length: int, inp: T, name: Optional[str] = None, final_view: Optional[View] = None) -> T:
        """UNet_Segmentation.py:476-503.  The last unit writes into `final_view` (a slice of the decoder concat)."""
        e = self.e
        name = name or self.anon("rp")
        lay = Layout.simple(filters)
        hw = inp.h * inp.w
        count = e.N * hw
        x = inp
        # the decoder level of the same resolution is the only consumer: the whole path is a side lane (Engine.cur_lane)
        # that overlaps the deeper levels in both passes
        self._lanes = getattr(self, "_lanes", 0) + 1
        lane = self._lanes if (final_view is not None and os.environ.get("SEMB_NO_LANES") is None) else 0
        prev_lane, e.cur_lane = e.cur_lane, lane
        try:
            x = self._res_path_units(filters, length, inp, name, final_view, lay, hw, count)
        finally:
            e.cur_lane = prev_lane
        x.lane = lane
        self.taps[name] = x
        return x

    def _res_path_units(self, filters, length, inp, name, final_view, lay, hw, count):
        e = self.e
        x = inp
        for i in range(length):
            s_raw, bns, o_raw, bno = self.res_unit_raw(x, lay)
            kact1 = self.kg.layer("activation", [o_raw.klayer])
            kadd = self.kg.layer("add", [s_raw.klayer, kact1])
            kact2 = self.kg.layer("activation", [kadd])
            bng, wl = self._bn(lay, True, count)
            kbn = self.kg.layer(bng.name, [kact2], wl)
            summed = e.new_buf(inp.h, inp.w, lay.phys, f"{name}_{i}_sum")
            e.add_op(AffineOp(e, hw, s_raw.view, bns, o_raw.view, bno, summed.view(), L.ACT_RELU, actb=L.ACT_RELU,
                              stats_out=bng.stats_ref()))
            e.add_op(bng)
            if i == length - 1 and final_view is not None:
                out_view = final_view
            else:
                out_view = e.new_buf(inp.h, inp.w, lay.phys, f"{name}_{i}").view()
            e.add_op(AffineOp(e, hw, summed.view(), bng, None, None, out_view, L.ACT_NONE))
            x = T(out_view, inp.h, inp.w, lay, kbn)
        return x

    def up_concat(self, x: T, filters: int, skip_buf: Buf, skip: T, name: str) -> T:
        """concatenate([Conv2DTranspose(filters,(2,2),strides=2)(x), skip])  (:542-551).
        `skip` already lives in channels [pad8(filters):) of skip_buf."""
        e = self.e
        self.nct += 1
        lay = Layout.simple(filters)
        base = f"conv2d_transpose_{self.nct}"
        wname, bname = base + "/kernel", base + "/bias"
        # The Keras kernel (kh,kw,Cout,Cin) is stored as the HWIO kernel of a 1x1 conv with 4*Cout outputs, one block
        # of Cout per kernel position: phys[0,0,ci,(2r+s)*Cout_p+co] = keras[r,s,co,ci].  The transposed conv is then
        # that 1x1 conv (tensor cores) followed by a depth-to-space pass that writes into the skip-concat buffer.
        cin_l, cin_p, co_p = x.layout.logical, x.layout.phys, lay.phys
        imap, omap = x.layout.index_map(), lay.index_map()

        def to_phys(arr, cin_p=cin_p, co_p=co_p, imap=imap, omap=omap):
            out = np.zeros((1, 1, cin_p, 4 * co_p), dtype=np.float32)
            for r in range(2):
                for s_ in range(2):
                    blk = out[0, 0, :, (2 * r + s_) * co_p:(2 * r + s_ + 1) * co_p]
                    blk[np.ix_(imap, omap)] = arr[r, s_].T            # (Cout,Cin) -> (Cin,Cout)
            return out

        def to_logical(phys, co_p=co_p, imap=imap, omap=omap, filters=filters, cin_l=cin_l):
            out = np.zeros((2, 2, filters, cin_l), dtype=np.float32)
            for r in range(2):
                for s_ in range(2):
                    blk = phys[0, 0, :, (2 * r + s_) * co_p:(2 * r + s_ + 1) * co_p]
                    out[r, s_] = blk[np.ix_(imap, omap)].T
            return out

        self._add(ParamSpec(wname, "convT_kernel", (2, 2, filters, cin_l), (1, 1, cin_p, 4 * co_p), {}, True, "glorot",
                            (4 * filters, 4 * cin_l), to_phys_fn=to_phys, to_logical_fn=to_logical))
        self._add(ParamSpec(bname, "convT_bias", (filters,), (lay.phys,), {0: lay.index_map()}, True, "zeros"))
        kct = self.kg.layer(base, [x.klayer], [wname, bname])
        kcat = self.kg.layer("concatenate", [kct, skip.klayer])
        up_view = skip_buf.view(0, lay.phys)
        y4 = e.new_buf(x.h, x.w, 4 * co_p, name + "_y4")
        ct = ConvOp(e, x.view, y4.view(), (x.h, x.w), (x.h, x.w), wname, None, 1, 1, (0, 0), L.PAD_ZERO, False)
        e.add_op(ct)
        e.add_op(ShuffleOp(e, y4.view(), up_view, x.h, x.w, bname, conv=ct))
        e.join_next(getattr(skip, "lane", 0))       # the next op (the decoder block's first conv) reads the res_path's output
        return T(skip_buf.view(), 2 * x.h, 2 * x.w, Layout.concat(lay, skip.layout), kcat)

    def _build(self, in_channels: int):
        e, f = self.e, self.filters
        H, W = self.h, self.w
        ph, pw = (16 - H % 16) % 16, (16 - W % 16) % 16
        lin = Layout.simple(in_channels)
        self.in_buf = e.new_buf(H, W, lin.phys, "input", requires_grad=False)
        kin = self.kg.layer("input", [])
        x = T(self.in_buf.view(), H, W, lin, kin)
        kpad = self.kg.layer("reflection_padding2d", [kin])
        if ph or pw:
            padded = e.new_buf(H + ph, W + pw, lin.phys, "input_padded", requires_grad=False)
            e.add_op(PadCropOp(e, x.view, padded.view(), (H, W), (H + ph, W + pw), ph // 2, pw // 2, "reflect"))
            x = T(padded.view(), H + ph, W + pw, lin, kpad)
        else:
            x = T(x.view, H, W, lin, kpad)
        PH, PW = H + ph, W + pw

        def pool(t: T) -> T:
            k = self.kg.layer("max_pooling2d", [t.klayer])
            out = e.new_buf(t.h // 2, t.w // 2, t.layout.phys, f"pool_{t.h // 2}")
            e.add_op(PoolOp(e, t.view, out.view(), t.h, t.w))
            return T(out.view(), t.h // 2, t.w // 2, t.layout, k)

        fp = pad8
        m1 = self.multi_res_block(f, x, "mres1")
        p1 = pool(m1)
        cat9 = e.new_buf(PH, PW, fp(f) * 2, "up9")
        r1 = self.res_path(f, 4, m1, "rp1", cat9.view(fp(f), fp(f)))
        m2 = self.multi_res_block(f * 2, p1, "mres2")
        p2 = pool(m2)
        cat8 = e.new_buf(PH // 2, PW // 2, fp(f * 2) * 2, "up8")
        r2 = self.res_path(f * 2, 3, m2, "rp2", cat8.view(fp(f * 2), fp(f * 2)))
        m3 = self.multi_res_block(f * 4, p2, "mres3")
        p3 = pool(m3)
        cat7 = e.new_buf(PH // 4, PW // 4, fp(f * 4) * 2, "up7")
        r3 = self.res_path(f * 4, 2, m3, "rp3", cat7.view(fp(f * 4), fp(f * 4)))
        m4 = self.multi_res_block(f * 8, p3, "mres4")
        p4 = pool(m4)
        cat6 = e.new_buf(PH // 8, PW // 8, fp(f * 8) * 2, "up6")
        r4 = self.res_path(f * 8, 1, m4, "rp4", cat6.view(fp(f * 8), fp(f * 8)))
        m5 = self.multi_res_block(f * 16, p4, "mres5")

        u6 = self.up_concat(m5, f * 8, cat6, r4, "up6")
        m6 = self.multi_res_block(32 * 8, u6, "mres6")     # decoder widths hard-coded 32*k (:543-549)
        u7 = self.up_concat(m6, f * 4, cat7, r3, "up7")
        m7 = self.multi_res_block(32 * 4, u7, "mres7")
        u8 = self.up_concat(m7, f * 2, cat8, r2, "up8")
        m8 = self.multi_res_block(32 * 2, u8, "mres8")
        u9 = self.up_concat(m8, f, cat9, r1, "up9")
        m9 = self.multi_res_block(f, u9, "mres9")

        kcrop = self.kg.layer("cropping2d", [m9.klayer])
        if ph or pw:
            cropped = e.new_buf(H, W, m9.layout.phys, "cropped")
            e.add_op(PadCropOp(e, m9.view, cropped.view(), (PH, PW), (H, W), ph // 2, pw // 2, "crop"))
            c = T(cropped.view(), H, W, m9.layout, kcrop)
        else:
            c = T(m9.view, H, W, m9.layout, kcrop)
        lout = Layout.simple(1)
        head_raw, bnh = self.conv2d_bn_raw(c, lout, 1)
        self.head_raw, self.head_norm = head_raw, bnh      # the loss recomputes the probability from the pre-activation
        kout = self.kg.layer("output", [head_raw.klayer])
        self.out_buf = e.new_buf(H, W, lout.phys, "output")
        e.add_op(AffineOp(e, H * W, head_raw.view, bnh, None, None, self.out_buf.view(), L.ACT_SIGMOID))
        self.output = T(self.out_buf.view(), H, W, lout, kout)
        self.keras_output_layer = kout

    def keras_weight_names(self) -> List[str]:
        return self.kg.weight_names(self.keras_output_layer)
