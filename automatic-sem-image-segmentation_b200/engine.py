"""Static-graph executor over libsemb200.so.

A network is a list of ops recorded once for a fixed (N, H, W); `forward()` replays the list,
`backward()` replays it in reverse calling each op's hand-written gradient.  torch is used only to
own device memory, streams and CUDA graphs; all arithmetic happens in the C-ABI kernels.

Channel padding: every activation buffer stores its channels in a *physical* layout whose segments
are padded to multiples of 8 (see include/semb200.h).  `Layout` keeps the logical<->physical map;
parameters are stored physically (zero in the padded rows/columns) in flat fp32 buffers so that Adam
and the NCCL all-reduce each run over ONE contiguous tensor per network.
"""
from __future__ import annotations

import ctypes as C
import os as _os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def pad8(c: int) -> int:
    return (c + 7) // 8 * 8


class Layout:
    """Channel layout: list of (logical_len, physical_len) segments."""

    def __init__(self, segs: Sequence[Tuple[int, int]]):
        self.segs = [(int(l), int(p)) for l, p in segs]
        self.logical = sum(l for l, _ in self.segs)
        self.phys = sum(p for _, p in self.segs)

    @staticmethod
    def simple(c: int) -> "Layout":
        return Layout([(c, pad8(c))])

    @staticmethod
    def concat(*layouts: "Layout") -> "Layout":
        segs = []
        for l in layouts:
            segs.extend(l.segs)
        return Layout(segs)

    def index_map(self) -> np.ndarray:
        """physical index of every logical channel"""
        out, base = [], 0
        for l, p in self.segs:
            out.extend(range(base, base + l))
            base += p
        return np.asarray(out, dtype=np.int64)


class View:
    """Channel slice of a Buf (what the C ABI calls semb_tensor)."""

    def __init__(self, buf: "Buf", coff: int, c: int, layout: Optional[Layout] = None):
        assert coff % 8 == 0 and c % 8 == 0 and coff + c <= buf.pitch, (coff, c, buf.pitch)
        self.buf, self.coff, self.C = buf, coff, c
        self.layout = layout
        self._t = None
        self._g = None

    @property
    def t(self) -> L.Tensor:
        if self._t is None:
            self._t = L.Tensor(self.buf.data.data_ptr(), self.C, self.buf.pitch, self.coff)
        return self._t

    @property
    def g(self) -> L.Tensor:
        if self._g is None:
            self._g = L.Tensor(self.buf.grad_tensor().data_ptr(), self.C, self.buf.pitch, self.coff)
        return self._g

    @property
    def requires_grad(self):
        return self.buf.requires_grad

    def torch_view(self, grad: bool = False) -> torch.Tensor:
        t = self.buf.grad_tensor() if grad else self.buf.data
        return t[..., self.coff:self.coff + self.C]


class Buf:
    def __init__(self, eng: "Engine", n: int, h: int, w: int, pitch: int, name: str, requires_grad: bool = True):
        assert pitch % 8 == 0
        self.eng, self.N, self.H, self.W, self.pitch, self.name = eng, n, h, w, pitch, name
        self.requires_grad = requires_grad
        self.data = torch.zeros((n, h, w, pitch), dtype=eng.tdtype, device=eng.device)
        self._grad = None
        self.grad_cover: List[Tuple[int, int]] = []  # channel intervals already written in this backward plan
        self.force_acc = False   # gradient buffer shared between engines: always accumulate, the owner zeroes it

    def grad_tensor(self) -> torch.Tensor:
        if self._grad is None:
            self._grad = torch.zeros_like(self.data)
        return self._grad

    def view(self, coff: int = 0, c: Optional[int] = None, layout: Optional[Layout] = None) -> View:
        return View(self, coff, self.pitch - coff if c is None else c, layout)

    def alias(self, n: int, h: int, w: int, pitch: int, name: str) -> "Buf":
        """The same storage (values AND gradient) seen with another NHWC shape: keras Flatten / Reshape of a dense NHWC tensor
        (WassersteinGAN.py:617,657).  All gradient writers must go through ONE of the two objects (the plans are per object)."""
        assert n * h * w * pitch == self.data.numel(), (self.name, (n, h, w, pitch), tuple(self.data.shape))
        return _AliasBuf(self, n, h, w, pitch, name)


class _AliasBuf(Buf):
    def __init__(self, base: Buf, n: int, h: int, w: int, pitch: int, name: str):
        self.eng, self.N, self.H, self.W, self.pitch, self.name = base.eng, n, h, w, pitch, name
        self.base = base
        self.requires_grad = base.requires_grad
        self.data = base.data.view(n, h, w, pitch)
        self._grad = None
        self.grad_cover = []
        self.force_acc = False

    def grad_tensor(self) -> torch.Tensor:
        if self._grad is None:
            self._grad = self.base.grad_tensor().view(self.N, self.H, self.W, self.pitch)
        return self._grad


class FlatStore:
    """Named fp32 slices of one flat device tensor (parameters, optimizer state, scratch)."""

    def __init__(self):
        self.entries: Dict[str, Tuple[int, int]] = {}
        self.order: List[str] = []
        self.size = 0
        self.t: Optional[torch.Tensor] = None

    def add(self, name: str, n: int) -> str:
        assert name not in self.entries, name
        n4 = (n + 3) // 4 * 4
        self.entries[name] = (self.size, n)
        self.order.append(name)
        self.size += n4
        return name

    def alloc(self, device):
        self.t = torch.zeros(max(self.size, 4), dtype=torch.float32, device=device)

    def ptr(self, name: str, extra: int = 0) -> int:
        return self.t.data_ptr() + 4 * (self.entries[name][0] + extra)

    def get(self, name: str) -> torch.Tensor:
        o, n = self.entries[name]
        return self.t[o:o + n]


class ParamSpec:
    """One Keras variable: logical shape + how its channel axes map into the physical tensor."""

    def __init__(self, name: str, kind: str, logical_shape, phys_shape, axis_maps: Dict[int, np.ndarray], trainable: bool,
                 init: str = "zeros", fans: Tuple[int, int] = (1, 1), to_phys_fn=None, to_logical_fn=None):
        self.name, self.kind = name, kind
        self.logical_shape, self.phys_shape = tuple(logical_shape), tuple(phys_shape)
        self.axis_maps, self.trainable, self.init, self.fans = axis_maps, trainable, init, fans
        self.to_phys_fn, self.to_logical_fn = to_phys_fn, to_logical_fn

    def to_phys(self, arr: np.ndarray) -> np.ndarray:
        arr = np.asarray(arr, dtype=np.float32)
        assert tuple(arr.shape) == self.logical_shape, (self.name, arr.shape, self.logical_shape)
        if self.to_phys_fn is not None:
            out = np.asarray(self.to_phys_fn(arr), dtype=np.float32)
            assert tuple(out.shape) == self.phys_shape, (self.name, out.shape, self.phys_shape)
            return out
        out = np.zeros(self.phys_shape, dtype=np.float32)
        idx = [np.arange(s) for s in self.logical_shape]
        for ax, m in self.axis_maps.items():
            idx[ax] = m
        out[np.ix_(*idx)] = arr
        return out

    def to_logical(self, phys: np.ndarray) -> np.ndarray:
        if self.to_logical_fn is not None:
            return np.ascontiguousarray(self.to_logical_fn(phys.reshape(self.phys_shape)))
        idx = [np.arange(s) for s in self.logical_shape]
        for ax, m in self.axis_maps.items():
            idx[ax] = m
        return np.ascontiguousarray(phys.reshape(self.phys_shape)[np.ix_(*idx)])


class Engine:
    """Owns buffers, parameters, scratch and the op list of ONE network instance."""

    def __init__(self, n: int, dtype: str = "bf16", device: Optional[torch.device] = None, dry: bool = False,
                 use_tc: bool = True, share: Optional["Engine"] = None):
        # "f32tc": fp32 storage with the stride-1 convs on the tensor cores (split bf16 operands); "f32": CUDA-core fp32
        split = dtype == "f32tc" or (dtype == "f32" and _os.environ.get("SEMB_F32_TC") == "1")
        if dtype == "f32tc":
            dtype = "f32"
        # dry=True builds the op list / parameter maps on the CPU for host-logic tests; nothing can execute.
        self.dry = dry
        if not dry:
            L.require_device()
        self.lib = L.load()
        self.N = n
        self.dtype_name = dtype
        self.dtype = L.BF16 if dtype == "bf16" else L.F32
        self.tdtype = torch.bfloat16 if dtype == "bf16" else torch.float32
        self.device = device or (torch.device("cpu") if dry else torch.device("cuda", torch.cuda.current_device()))
        self.ops: List["Op"] = []
        self.params = FlatStore()        # trainable
        self.state = FlatStore()         # moving statistics
        self.zeroed = FlatStore()        # scratch that must be zero at step start (moments, bwd sums, loss sums)
        self.scratch = FlatStore()       # scale/shift/mean/invstd/c1/c2
        self.specs: Dict[str, ParamSpec] = {}
        self.spec_order: List[str] = []
        self.bufs: List[Buf] = []
        self.finalized = False
        self.grads = self.adam_m = self.adam_v = None
        self._keep = []
        # tensor-core path: bf16 storage only; packed bf16 weight images are refreshed after every weight change
        self.tc_enabled = bool(use_tc) and dtype == "bf16"
        # fp32 storage on the tensor cores (dtype "f32tc"): stride-1 zero-padded convs run the same tcgen05 kernels on
        # split bf16 operands with fp32 results (semb_split_bf16 / semb_conv2d_fwd_tc_f32).  Measured on the reference's
        # dataset: sigmoid maps within 1e-3 (max 8e-4) of the oracle, 0-1 mask pixels of 720 896 differ -- the tensor
        # core's fp32 accumulation is not round-to-nearest -- so the strict parity mode ("f32": bit-exact masks) stays on
        # the CUDA-core kernels
        self.tc_split = bool(use_tc) and split
        self.split_terms = 3 if _os.environ.get("SEMB_SPLIT_TERMS") == "3" else 6      # 3: 2^-16 products, 6: fp32-grade
        self._split_need = [0, 0]        # bf16 elements of the two shared split-operand scratch regions (x / dy)
        self._split_buf = [None, None]
        self.tc_packs: List[dict] = []
        self._pack_dirty = True
        self._pack_table = None
        # one cooperative launch for the BN backward (sums -> grid barrier -> gradients).  Measured on B200 (round 1): SLOWER
        # than the two-pass kernels on every layer of the UNet (5.41 vs 4.77 ms per step; the barrier and the cooperative
        # launch cost more than the L2 re-read saves), so it stays opt-in.
        self.fused_affine_bwd = _os.environ.get("SEMB_FUSED_AFFINE_BWD") is not None
        # the C-length finalize kernels (scale/shift from moments; c1/c2/dgamma/dbeta from the backward sums) are folded
        # into the affine kernels that consume them: ~170 launches of 5-6 us fewer per UNet step
        self.fold_norm = _os.environ.get("SEMB_NO_FOLD_NORM") is None
        # weight gradients run on a side stream, concurrently with the data-gradient / normalisation chain (they only
        # meet again at the optimizer); set by the model front end, off for weight-sharing towers
        self.wgrad_stream: Optional[torch.cuda.Stream] = None
        self.skip_wgrad = False
        # Side lanes: groups of ops whose results the main chain needs only much later (the res_path of every encoder level
        # is consumed by the decoder level of the same resolution) are issued on their own stream and so overlap the deep,
        # low-resolution part of the network, whose 64-1024-tile kernels leave most SMs idle.  Builders tag ops through
        # `cur_lane` / `join_next`; the model front end switches the streams on (`enable_lanes`).
        self.cur_lane = 0
        self._pending_join: set = set()
        self.lane_streams: Dict[int, torch.cuda.Stream] = {}
        self.lanes_on = False
        self.bwd_order: Optional[List["Op"]] = None
        # weight sharing between towers of the same network (CycleGAN applies each generator three times per step):
        # a sharing engine has its own buffers / ops / scratch but uses the root's parameters, gradients and packs
        self.share = share
        # stride-2 convs on the stride-1 tensor-core kernels: virtual (3,3,4*Cin,Cout) kernels over the space-to-depth image
        self.s2d: Dict[tuple, dict] = {}
        if share is not None:
            assert share.finalized and share.dtype == self.dtype
            self.params, self.state, self.specs, self.spec_order = share.params, share.state, share.specs, share.spec_order
            self.tc_packs = share.tc_packs
            self.s2d = share.s2d

    # ---- construction ------------------------------------------------------------------
    def new_buf(self, h: int, w: int, pitch: int, name: str, requires_grad: bool = True, n: Optional[int] = None) -> Buf:
        b = Buf(self, self.N if n is None else n, h, w, pitch, name, requires_grad)
        self.bufs.append(b)
        return b

    def add_param(self, spec: ParamSpec):
        if self.share is not None:
            assert spec.name in self.specs and self.specs[spec.name].phys_shape == spec.phys_shape, spec.name
            return
        self.specs[spec.name] = spec
        self.spec_order.append(spec.name)
        store = self.params if spec.trainable else self.state
        store.add(spec.name, int(np.prod(spec.phys_shape)))

    def add_op(self, op: "Op"):
        op.lane = self.cur_lane
        op.join_fwd, op.join_bwd = set(), set()
        if self.cur_lane == 0 and self._pending_join:
            op.join_fwd, self._pending_join = self._pending_join, set()     # first main-lane consumer of the lanes' results
        self.ops.append(op)
        return op

    def join_next(self, lane: int):
        """The next op added on the main lane reads what `lane` produces."""
        if lane:
            self._pending_join.add(lane)

    def enable_lanes(self):
        if any(op.lane for op in self.ops) and not self.dry:
            for l in sorted({op.lane for op in self.ops if op.lane}):
                self.lane_streams[l] = torch.cuda.Stream(device=self.device)
            self.lanes_on = True

    def _plan_bwd_order(self):
        """Backward issue order: reversed forward order, except that the ops of a side lane move up to right after the
        backward of the op that joined the lane in the forward pass -- the earliest reader of the lane's output, hence the
        op that completes its gradient -- instead of waiting behind the whole deep part of the network.  The op that
        followed the group in plain reversed order (it accumulates into the lane's input gradient) joins the lane."""
        rev = list(reversed(self.ops))
        lanes = sorted({op.lane for op in self.ops if op.lane})
        for l in lanes:
            group = [op for op in rev if op.lane == l]
            first = rev.index(group[0])
            after = [op for op in rev[first:] if op.lane == 0]
            joiner = [op for op in rev if l in op.join_fwd]
            if not joiner or rev.index(joiner[0]) > first:
                continue            # no consumer recorded (or it already comes later): keep the plain order
            if after:
                after[0].join_bwd.add(l)
            rest = [op for op in rev if op.lane != l]
            at = rest.index(joiner[0]) + 1
            rev = rest[:at] + group + rest[at:]
        self.bwd_order = rev

    def finalize(self):
        if self.share is not None:
            for s in (self.zeroed, self.scratch):
                s.alloc(self.device)
            r = self.share
            self.grads, self.adam_m, self.adam_v, self.adam_state, self.lr = r.grads, r.adam_m, r.adam_v, r.adam_state, r.lr
            for op in reversed(self.ops):
                op.plan_backward()
            self._plan_bwd_order()
            self.finalized = True
            return
        for s in (self.params, self.state, self.zeroed, self.scratch):
            s.alloc(self.device)
        self.grads = torch.zeros_like(self.params.t)
        self.adam_m = torch.zeros_like(self.params.t)
        self.adam_v = torch.zeros_like(self.params.t)
        self.adam_state = torch.zeros(4, dtype=torch.int32, device=self.device)   # {int64 t; float alpha; float pad}
        self.lr = torch.zeros(1, dtype=torch.float32, device=self.device)
        # static plan of gradient accumulation: walk the ops in backward order once
        for op in reversed(self.ops):
            op.plan_backward()
        self._plan_bwd_order()
        for pk in self.tc_packs:
            nbytes = int(self.lib.semb_pack_weights_tc(None, pk["R"], pk["S"], pk["Cin"], pk["Cout"], pk["flip"], None, None))
            if nbytes < 0:
                L.check(nbytes)
            pk["buf"] = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device=self.device)
        recs = [r for r in self.s2d.values() if r["dw3"] is not None]
        if recs and not self.dry:
            flat = torch.zeros(sum(r["dw3"].numel() for r in recs), dtype=torch.float32, device=self.device)
            o = 0
            for r in recs:
                n = r["dw3"].numel()
                r["dw3"], r["flat"] = flat[o:o + n], True
                o += n
            self._vgrad_flat = flat
        self.finalized = True

    def s2d_weight(self, w: str, k: int, pt: int, pl: int, cin: int, cout: int) -> dict:
        """Virtual fp32 kernel w3 (3,3,4*cin,cout) of the stride-2 conv `w` and the buffer its gradient is reduced into."""
        root = self.share or self
        key = (w, pt, pl)
        rec = root.s2d.get(key)
        if rec is None:
            n = 9 * 4 * cin * cout
            rec = {"kind": "s2d", "w": w, "k": k, "pt": pt, "pl": pl, "cin": cin, "cout": cout, "key": f"{w}/s2d{pt}{pl}",
                   "w3": torch.zeros(n, dtype=torch.float32, device=self.device) if not self.dry else None,
                   "dw3": torch.zeros(n, dtype=torch.float32, device=self.device) if not self.dry else None}
            root.s2d[key] = rec
            root._pack_dirty = True
        return rec

    def merge_weight(self, wa: str, ws: str, cin: int, ca: int, cs: int) -> dict:
        """Virtual fp32 kernel w3 (3,3,cin,ca+cs) = [3x3 kernel `wa` | 1x1 kernel `ws` on the centre tap] (semb_merge_weights) and
        the buffer its gradient is reduced into before it is folded into the two Keras-layout gradients."""
        root = self.share or self
        key = (wa, "merge", ws)
        rec = root.s2d.get(key)
        if rec is None:
            n = 9 * cin * (ca + cs)
            rec = {"kind": "merge", "w": wa, "w2": ws, "k": 3, "cin": cin, "ca": ca, "cs": cs, "cout": ca + cs, "key": f"{wa}/merge",
                   "w3": torch.zeros(n, dtype=torch.float32, device=self.device) if not self.dry else None,
                   "dw3": torch.zeros(n, dtype=torch.float32, device=self.device) if not self.dry else None}
            root.s2d[key] = rec
            root._pack_dirty = True
        return rec

    def split3_weight(self, w: str, k: int, cin: int, cout: int, axis: int) -> dict:
        """Stacked fp32 kernel [w ; w ; w - bf16(w)] of `w` along the input (axis 0) or output (axis 1) channels; the
        packer rounds it to the bf16 images the split-operand convs contract with (semb_split3_weights)."""
        root = self.share or self
        key = (w, "split3", axis)
        rec = root.s2d.get(key)
        if rec is None:
            n = k * k * self.split_terms * cin * cout
            rec = {"kind": "split3", "axis": axis, "w": w, "k": k, "cin": cin, "cout": cout, "key": f"{w}/split3_{axis}", "terms": self.split_terms,
                   "w3": torch.zeros(n, dtype=torch.float32, device=self.device) if not self.dry else None, "dw3": None}
            root.s2d[key] = rec
            root._pack_dirty = True
        return rec

    def split_scratch(self, which: int, c3: int, nelem: int) -> L.Tensor:
        """View (3C channels, bf16) over the shared split-operand scratch region `which` (0: conv inputs, 1: gradients)."""
        buf = self._split_buf[which]
        if buf is None or buf.numel() < nelem:
            buf = torch.empty(max(nelem, self._split_need[which]), dtype=torch.bfloat16, device=self.device)
            self._split_buf[which] = buf
        return L.Tensor(buf.data_ptr(), c3, c3, 0)

    def tapfold_weight(self, w: str, k: int, cin: int, cout: int, kind: str) -> dict:
        """Virtual 1x1 kernel of a k x k conv with ONE input channel ('stem': (T8, cout)) or ONE output channel
        ('head': (cin, T8)), T8 = pad8(k*k); see semb_tapfold_weights."""
        root = self.share or self
        key = (w, kind)
        rec = root.s2d.get(key)
        if rec is None:
            t8 = (k * k + 7) // 8 * 8
            n = t8 * (cout if kind == "stem" else cin)
            rec = {"kind": kind, "w": w, "k": k, "cin": cin, "cout": cout, "key": f"{w}/{kind}", "t8": t8,
                   "w3": torch.zeros(n, dtype=torch.float32, device=self.device) if not self.dry else None,
                   "dw3": torch.zeros(n, dtype=torch.float32, device=self.device) if not self.dry else None}
            root.s2d[key] = rec
            root._pack_dirty = True
        return rec

    def _virtual_weight_kernel(self, rec: dict, master_ptr: int, virt_ptr: int, direction: int):
        if rec["kind"] == "merge":
            second = self.gptr(rec["w2"]) if direction else self.params.ptr(rec["w2"])      # weights (0) or gradients (1) of the shortcut
            L.check(self.lib.semb_merge_weights(master_ptr, second, rec["cin"], rec["ca"], rec["cs"], virt_ptr, direction, self.stream))
        elif rec["kind"] == "split3":
            assert direction == 0       # gradients of the split-operand convs go straight to the master gradient
            L.check(self.lib.semb_split_weights(master_ptr, rec["k"], rec["k"], rec["cin"], rec["cout"], virt_ptr, rec["axis"], rec["terms"],
                                                self.stream))
        elif rec["kind"] == "s2d":
            L.check(self.lib.semb_s2d_weights(master_ptr, rec["k"], rec["pt"], rec["pl"], rec["cin"], rec["cout"], virt_ptr, direction,
                                              self.stream))
        else:
            L.check(self.lib.semb_tapfold_weights(master_ptr, rec["k"], rec["cin"], rec["cout"], virt_ptr,
                                                  0 if rec["kind"] == "stem" else 1, direction, self.stream))

    def zero_grads(self):
        """Zeroes the flat gradient buffer and the gradients of the virtual stride-2 kernels."""
        assert self.share is None
        st = self.stream
        L.check(self.lib.semb_fill_f32(self.grads.data_ptr(), self.grads.numel(), 0.0, st))
        self._zero_virtual_grads()
        self._vgrads_folded = False

    def _zero_virtual_grads(self):
        """The gradients of the virtual kernels live in ONE flat buffer (one fill); records created after finalize() keep
        their own tensors."""
        st = self.stream
        flat = self.__dict__.get("_vgrad_flat")
        if flat is not None:
            L.check(self.lib.semb_fill_f32(flat.data_ptr(), flat.numel(), 0.0, st))
        for rec in self.s2d.values():
            if rec["dw3"] is not None and not rec.get("flat"):
                L.check(self.lib.semb_fill_f32(rec["dw3"].data_ptr(), rec["dw3"].numel(), 0.0, st))

    def fold_virtual_grads(self):
        """Adds the gradients of the virtual stride-2 kernels into the Keras-layout gradients (before all-reduce / Adam)."""
        assert self.share is None
        st = self.stream
        for rec in self.s2d.values():
            if rec["dw3"] is None:
                continue
            self._virtual_weight_kernel(rec, self.gptr(rec["w"]), rec["dw3"].data_ptr(), 1)
        self._zero_virtual_grads()
        self._vgrads_folded = True

    def tc_pack(self, w: str, R: int, S: int, cin: int, cout: int, flip: int, vw: Optional[dict] = None) -> dict:
        for pk in self.tc_packs:
            if pk["w"] == w and pk["flip"] == flip:
                return pk
        pk = {"w": w, "R": R, "S": S, "Cin": cin, "Cout": cout, "flip": flip, "buf": None, "vw": vw}
        self.tc_packs.append(pk)
        root = self.share or self
        if root.finalized:      # a tower added after the root was finalised needs a pack the root never used
            nbytes = int(self.lib.semb_pack_weights_tc(None, R, S, cin, cout, flip, None, None))
            pk["buf"] = torch.zeros(nbytes // 2, dtype=torch.bfloat16, device=self.device)
            root._pack_dirty = True
        return pk

    def _pack_jobs(self):
        """Device table of pack jobs (one launch re-packs every weight image); rebuilt when a tower adds a pack."""
        if self._pack_table is None or self._pack_table[1] != len(self.tc_packs):
            n = len(self.tc_packs)
            host = np.zeros(n * int(self.lib.semb_pack_batch_job_size()), dtype=np.uint8)
            blocks = 0
            for i, pk in enumerate(self.tc_packs):
                wptr = pk["vw"]["w3"].data_ptr() if pk.get("vw") else self.params.ptr(pk["w"])
                blocks = int(self.lib.semb_pack_batch_prepare(i, wptr, pk["buf"].data_ptr(), pk["R"], pk["S"],
                                                               pk["Cin"], pk["Cout"], pk["flip"], blocks, host.ctypes.data))
                if blocks < 0:
                    L.check(blocks)
            self._pack_table = (torch.from_numpy(host).to(self.device), n, blocks)
        return self._pack_table

    def repack(self):
        """fp32 master weights -> packed bf16 UMMA images (after set_weights / Adam)."""
        if self.share is not None:
            return self.share.repack()
        if not self.dry and self.tc_packs:
            for rec in self.s2d.values():       # virtual (space-to-depth / tap-folded) kernels follow the master weights first
                self._virtual_weight_kernel(rec, self.params.ptr(rec["w"]), rec["w3"].data_ptr(), 0)
            table, n, blocks = self._pack_jobs()
            L.check(self.lib.semb_pack_weights_tc_batch(table.data_ptr(), n, blocks, self.stream))
        self._pack_dirty = False

    def wgrad_tc(self, geom, x_t, dy_t, dw_ptr: int):
        """Tensor-core weight gradient; layers with a planar path (semb_conv2d_wgrad_tc_workspace > 0) get the network's
        shared scratch (weight gradients of one network are serialised on one stream)."""
        key = (geom.N, geom.H, geom.W, geom.OH, geom.OW, geom.Cin, geom.Cout, geom.R, geom.pad_t, geom.pad_l, geom.pad_mode)
        root = self.share or self
        cache = root.__dict__.setdefault("_wgrad_ws_need", {})
        need = cache.get(key)
        if need is None:
            need = int(self.lib.semb_conv2d_wgrad_tc_workspace(C.byref(geom))) if _os.environ.get("SEMB_WGRAD_NO_PLANAR") is None else 0
            cache[key] = need
        if need <= 0:
            return L.check(self.lib.semb_conv2d_wgrad_tc(C.byref(geom), C.byref(x_t), C.byref(dy_t), dw_ptr, self.stream))
        ws = root.__dict__.get("_wgrad_ws")
        if ws is None or ws.numel() < need:
            ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
            root._wgrad_ws = ws
        base = (ws.data_ptr() + 127) // 128 * 128
        L.check(self.lib.semb_conv2d_wgrad_tc_ws(C.byref(geom), C.byref(x_t), C.byref(dy_t), dw_ptr, base, need, self.stream))

    def gptr(self, name: str) -> int:
        o, _ = self.params.entries[name]
        return self.grads.data_ptr() + 4 * o

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- weights -------------------------------------------------------------------------
    def set_param(self, name: str, arr: np.ndarray):
        spec = self.specs[name]
        store = self.params if spec.trainable else self.state
        store.get(name).copy_(torch.from_numpy(spec.to_phys(arr).reshape(-1)))
        self._pack_dirty = True

    def get_param(self, name: str) -> np.ndarray:
        spec = self.specs[name]
        store = self.params if spec.trainable else self.state
        return spec.to_logical(store.get(name).detach().cpu().numpy())

    def get_grad(self, name: str) -> np.ndarray:
        spec = self.specs[name]
        o, n = self.params.entries[name]
        return spec.to_logical(self.grads[o:o + n].detach().cpu().numpy())

    def init_params(self, seed: int = 0):
        """Glorot-uniform kernels, zeros / ones elsewhere, drawn in creation order from one generator."""
        gen = torch.Generator().manual_seed(seed)
        for name in self.spec_order:
            spec = self.specs[name]
            if spec.init == "glorot":
                fi, fo = spec.fans
                limit = float(np.sqrt(6.0 / (fi + fo)))
                arr = ((torch.rand(spec.logical_shape, generator=gen) * 2.0 - 1.0) * limit).numpy()
            elif spec.init == "ones":
                arr = np.ones(spec.logical_shape, dtype=np.float32)
            else:
                arr = np.zeros(spec.logical_shape, dtype=np.float32)
            self.set_param(name, arr)

    # ---- execution -------------------------------------------------------------------------
    def zero_step(self, zero_grads: bool):
        st = self.stream
        L.check(self.lib.semb_fill_f32(self.zeroed.t.data_ptr(), self.zeroed.t.numel(), 0.0, st))
        if zero_grads and self.share is None:
            self.zero_grads()

    def forward(self, training: bool):
        if self.dry:
            raise L.SembError("dry engine: kernels need an sm_100a device, there is no CPU execution path")
        root = self.share or self
        if root._pack_dirty and root.tc_packs:
            root.repack()
        self._run(self.ops, lambda op: op.fwd(training), "join_fwd")

    def _run(self, order, call, join_attr: str):
        """Issues `order` with side-lane ops on their streams: a lane forks from the main stream at its first op (it then
        depends on everything queued on the main stream so far) and is joined by the main-lane ops that name it."""
        if not self.lanes_on:
            for op in order:
                call(op)
            return
        main = torch.cuda.current_stream(self.device)
        active = set()
        for op in order:
            if op.lane:
                st = self.lane_streams[op.lane]
                if op.lane not in active:
                    st.wait_stream(main)
                    active.add(op.lane)
                with torch.cuda.stream(st):
                    call(op)
            else:
                for l in getattr(op, join_attr):
                    if l in active:
                        main.wait_stream(self.lane_streams[l])
                        active.discard(l)
                call(op)
        for l in active:
            main.wait_stream(self.lane_streams[l])

    def backward(self):
        (self.share or self)._vgrads_folded = False      # new gradients of the virtual kernels are about to be produced
        self._run(self.bwd_order if self.bwd_order is not None else list(reversed(self.ops)), lambda op: op.bwd(), "join_bwd")
        if self.wgrad_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.wgrad_stream)       # join before the optimizer

    def on_wgrad_stream(self, fn):
        """Runs `fn` (kernel launches that only produce weight gradients) on the side stream, after everything queued
        so far on the current stream; the caller's stream does not wait for it (see backward())."""
        if self.skip_wgrad:
            return None         # this tower only back-propagates data gradients (discriminator inside the generator phase)
        side = self.wgrad_stream
        if side is None:
            return fn()
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            fn()

    def adam(self, beta1: float, beta2: float, eps: float, gscale: float = 1.0):
        if self.s2d and not self.__dict__.get("_vgrads_folded", False):
            self.fold_virtual_grads()       # skipped when the caller already folded (before its all-reduce)
        L.check(self.lib.semb_adam_step(self.params.t.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                                        self.adam_v.data_ptr(), self.params.t.numel(), self.lr.data_ptr(),
                                        beta1, beta2, eps, gscale, self.adam_state.data_ptr(), self.stream))
        if self.tc_packs:
            self.repack()


def plan_grad_write(view: View) -> int:
    """Returns the accumulate flag for a gradient write into `view` at this point of the backward plan."""
    if view.buf.force_acc:
        return 1
    lo, hi = view.coff, view.coff + view.C
    cover = view.buf.grad_cover
    inside = any(a <= lo and hi <= b for a, b in cover)
    if inside:
        return 1
    overlap = any(not (hi <= a or b <= lo) for a, b in cover)
    if overlap:
        # a wider write after narrower ones: the kernels would overwrite the earlier slices
        raise RuntimeError(f"partial gradient overlap on {view.buf.name}: {cover} vs ({lo},{hi})")
    cover.append((lo, hi))
    return 0


class Op:
    eng: Engine

    def plan_backward(self):
        pass

    def fwd(self, training: bool):
        raise NotImplementedError

    def bwd(self):
        pass


# --------------------------------------------------------------------------------------------------
class ConvOp(Op):
    """Conv2D (transposed=False) or Conv2DTranspose (transposed=True).

    For transposed=True the geometry describes the equivalent strided conv that maps the OUTPUT of the
    transposed conv back to its INPUT (see semb_conv2d_dgrad in semb200.h); x is the small tensor."""

    def __init__(self, eng: Engine, x: View, y: View, hw_in: Tuple[int, int], hw_out: Tuple[int, int], w: str,
                 bias: Optional[str], k: int, stride: int, pad_tl: Tuple[int, int], pad_mode: int, transposed: bool,
                 stats: Optional[Tuple[str, int, int, int]] = None, n: Optional[int] = None):
        self.eng, self.x, self.y, self.w, self.bias, self.transposed = eng, x, y, w, bias, transposed
        n = eng.N if n is None else n
        if not transposed:
            h, wd = hw_in
            oh, ow = hw_out
            cin, cout = x.C, y.C
        else:
            oh, ow = hw_in     # the transposed conv's input is the strided conv's output
            h, wd = hw_out
            cin, cout = y.C, x.C
        self.geom = L.ConvGeom(n, h, wd, oh, ow, cin, cout, k, k, stride, pad_tl[0], pad_tl[1], pad_mode, eng.dtype)
        self.stats = stats  # (zeroed-store name, offset, nstride, cstride)
        self.acc_x = 0
        self.d2s_up = None  # set by ShuffleOp: (destination view, bias name, UH, UW) of a fused depth-to-space epilogue
        self.use_tc = (eng.tc_enabled and not transposed and stride == 1 and k in (1, 3))
        self.pad_buf = None
        if pad_mode == L.PAD_REFLECT and x.requires_grad and not transposed:
            # the data gradient of a reflect-padded conv is taken on the padded domain and folded back
            hp = max(h, (oh - 1) * stride + k)
            wp = max(wd, (ow - 1) * stride + k)
            self.pad_buf = eng.new_buf(hp, wp, x.C, f"{w}_dxpad", n=n)
            self.pad_hw = (hp, wp)
            self.geom_p = L.ConvGeom(n, hp, wp, oh, ow, cin, cout, k, k, stride, 0, 0, L.PAD_ZERO, eng.dtype)
        # Stride-2 3x3 / 4x4 convs (CycleGAN down / up-sampling, PatchGAN): 3x3-embedded 2x2 stride-1 conv over the
        # space-to-depth image of the big tensor, on the TMA / tcgen05 kernels (2.25x the FLOPs, but ~30x the CUDA-core rate).
        self.s2d = None
        # (k = 5 with a leading pad of 1 -- Keras 'same' on even sizes, the WGAN critic's convs -- fills the whole 3x3 kernel)
        if (eng.tc_enabled and stride == 2 and (k in (3, 4) or (k == 5 and tuple(pad_tl) == (1, 1) and not transposed))
                and pad_mode == L.PAD_ZERO and pad_tl[0] in (0, 1) and pad_tl[1] in (0, 1)
                and not (transposed and bias is not None and stats is not None) and _os.environ.get("SEMB_NO_S2D") is None):
            pt, pl = pad_tl
            h2, w2 = (h + 1) // 2, (wd + 1) // 2
            rec = eng.s2d_weight(w, k, pt, pl, cin, cout)
            big = y if transposed else x
            self.b2 = eng.new_buf(h2, w2, 4 * cin, f"{w}_s2d", requires_grad=big.requires_grad, n=n)
            self.s2d = rec
            self.s2d_hw = (h2, w2)
            self.geom_s = L.ConvGeom(n, h2, w2, oh, ow, 4 * cin, cout, 3, 3, 1, pt, pl, L.PAD_ZERO, eng.dtype)       # big' -> small
            self.geom_t = L.ConvGeom(n, oh, ow, h2, w2, cout, 4 * cin, 3, 3, 1, 2 - pt, 2 - pl, L.PAD_ZERO, eng.dtype)  # small -> big'
            self.pk_s = eng.tc_pack(rec["key"], 3, 3, 4 * cin, cout, 0, vw=rec)
            self.pk_t = eng.tc_pack(rec["key"], 3, 3, 4 * cin, cout, 1, vw=rec)
            self.stats4 = None
            if transposed and stats is not None:
                groups = n if stats[2] != 0 else 1
                self.stats4 = eng.zeroed.add(f"{w}/s2d_moments_{len(eng.ops)}", 2 * groups * 2 * 4 * cin)     # fp64: 2 floats each
                self.stats4_groups = groups
        # 7x7 reflect-padded convs with ONE input channel (generator stem) or ONE output channel (head): the 49 taps become
        # channels of a 1x1 tensor-core conv (im2col of a single channel / shift-and-add of a per-tap conv); on the CUDA
        # cores these two layers took 144 of the 215 ms of a CycleGAN step.
        self.tapfold = None
        self.tf_valid = False
        lshape = eng.specs[w].logical_shape if w in eng.specs else (k, k, cin, cout)       # LOGICAL channel counts decide
        # the PatchGAN output conv (4x4, stride 1, 'valid', ONE output channel; CycleGAN.py:448): same per-tap 1x1 conv +
        # shift-and-add as the generator head, on the un-padded input (on the CUDA cores it took 5 % of a CycleGAN step)
        valid_head = (k > 1 and stride == 1 and pad_mode == L.PAD_ZERO and tuple(pad_tl) == (0, 0) and not transposed and stats is None
                      and lshape[3] == 1 and lshape[2] > 1 and not transposed and tuple(hw_out) == (hw_in[0] - k + 1, hw_in[1] - k + 1))
        if (eng.tc_enabled and stride == 1 and not transposed and _os.environ.get("SEMB_NO_TAPFOLD") is None and
                (valid_head or (k == 7 and pad_mode == L.PAD_REFLECT and tuple(pad_tl) == (3, 3) and (lshape[2] == 1 or lshape[3] == 1)))):
            kind = "stem" if lshape[2] == 1 else "head"
            if not (kind == "head" and stats is not None):
                rec = eng.tapfold_weight(w, k, cin, cout, kind)
                t8 = rec["t8"]
                self.tapfold = rec
                self.tf_valid = valid_head
                # (th, tw) = OUTPUT size, (th+k-1, tw+k-1) = the (padded) input the per-tap conv runs over
                th, tw = (h - k + 1, wd - k + 1) if valid_head else (h, wd)
                self.tf_xpad = None if valid_head else eng.new_buf(th + k - 1, tw + k - 1, cin, f"{w}_xpad", requires_grad=x.requires_grad, n=n)
                if kind == "stem":
                    self.tf_mid = eng.new_buf(th, tw, t8, f"{w}_patches", requires_grad=x.requires_grad, n=n)
                    self.geom1 = L.ConvGeom(n, th, tw, th, tw, t8, cout, 1, 1, 1, 0, 0, L.PAD_ZERO, eng.dtype)
                    self.geom1d = L.ConvGeom(n, th, tw, th, tw, cout, t8, 1, 1, 1, 0, 0, L.PAD_ZERO, eng.dtype)
                    self.pk1 = eng.tc_pack(rec["key"], 1, 1, t8, cout, 0, vw=rec)
                    self.pk1d = eng.tc_pack(rec["key"], 1, 1, t8, cout, 1, vw=rec) if x.requires_grad else None
                else:
                    hp, wp = th + k - 1, tw + k - 1
                    self.tf_mid = eng.new_buf(hp, wp, t8, f"{w}_ztaps", requires_grad=True, n=n)
                    self.geom1 = L.ConvGeom(n, hp, wp, hp, wp, cin, t8, 1, 1, 1, 0, 0, L.PAD_ZERO, eng.dtype)
                    self.geom1d = L.ConvGeom(n, hp, wp, hp, wp, t8, cin, 1, 1, 1, 0, 0, L.PAD_ZERO, eng.dtype)
                    self.pk1 = eng.tc_pack(rec["key"], 1, 1, cin, t8, 0, vw=rec)
                    self.pk1d = eng.tc_pack(rec["key"], 1, 1, cin, t8, 1, vw=rec) if x.requires_grad else None
        # Reflect-padded tensor-core convs (CycleGAN residual blocks): the padded input is materialised once per forward
        # (a cheap copy next to a 512->512 conv) so that forward and weight gradient run the TMA kernels as zero-pad
        # "valid" convolutions over it; the data gradient already works on the padded domain (pad_buf above).
        self.x_pad = None
        if self.use_tc and pad_mode == L.PAD_REFLECT:
            hp = max(h, (oh - 1) * stride + k)
            wp = max(wd, (ow - 1) * stride + k)
            self.x_pad = eng.new_buf(hp, wp, x.C, f"{w}_xpad", requires_grad=False, n=n)
            self.x_pad_hw = (hp, wp)
            self.geom_v = L.ConvGeom(n, hp, wp, oh, ow, cin, cout, k, k, stride, 0, 0, L.PAD_ZERO, eng.dtype)
        # parity mode: fp32 tensors, tensor-core math on bf16 x 3 split operands (see Engine.tc_split)
        self.split = None
        if (eng.tc_split and not transposed and stride == 1 and k in (1, 3) and pad_mode == L.PAD_ZERO and bias is None
                and self.s2d is None and self.tapfold is None):
            r0 = eng.split3_weight(w, k, cin, cout, 0)
            T = eng.split_terms
            g3 = L.ConvGeom(n, h, wd, oh, ow, T * cin, cout, k, k, 1, pad_tl[0], pad_tl[1], L.PAD_ZERO, L.BF16)
            gw = L.ConvGeom(n, h, wd, oh, ow, cin, cout, k, k, 1, pad_tl[0], pad_tl[1], L.PAD_ZERO, L.BF16)
            self.split = {"g3": g3, "gw": gw, "pk": eng.tc_pack(r0["key"], k, k, T * cin, cout, 0, vw=r0), "nx": n * h * wd, "ny": n * oh * ow,
                          "T": T}
            eng._split_need[0] = max(eng._split_need[0], n * h * wd * T * cin)
            eng._split_need[1] = max(eng._split_need[1], n * oh * ow * T * cout)
            if x.requires_grad:
                r1 = eng.split3_weight(w, k, cin, cout, 1)
                self.split["gd"] = L.ConvGeom(n, oh, ow, h, wd, T * cout, cin, k, k, 1, k - 1 - pad_tl[0], k - 1 - pad_tl[1], L.PAD_ZERO, L.BF16)
                self.split["pkd"] = eng.tc_pack(r1["key"], k, k, cin, T * cout, 1, vw=r1)
        if self.use_tc:
            self.pk_fwd = eng.tc_pack(w, k, k, cin, cout, 0)
            self.pk_bwd = None
            if x.requires_grad:
                self.pk_bwd = eng.tc_pack(w, k, k, cin, cout, 1)
                # stride-1 data gradient = conv of dy with the mirrored, transposed kernel, pad k-1-pad
                if self.pad_buf is None:
                    self.geom_d = L.ConvGeom(n, oh, ow, h, wd, cout, cin, k, k, 1, k - 1 - pad_tl[0], k - 1 - pad_tl[1],
                                             L.PAD_ZERO, eng.dtype)
                else:
                    self.geom_d = L.ConvGeom(n, oh, ow, self.pad_hw[0], self.pad_hw[1], cout, cin, k, k, 1, k - 1, k - 1,
                                             L.PAD_ZERO, eng.dtype)
                    # (Tiling the batch as ONE tall image with zero separator rows -- 85 instead of 120 tiles for 8 x 34 x 34 -- was
                    # measured in round 2: no gain, 170 CTAs still need two waves of the 148 SMs; removed.)

    def plan_backward(self):
        if self.x.requires_grad:
            self.acc_x = plan_grad_write(self.x)

    def _stats_args(self):
        if self.stats is None:
            return None, 0, 0
        name, off, ns, cs = self.stats
        return self.eng.zeroed.ptr(name, 2 * off), ns, cs      # offsets/strides are in doubles

    def _shuffle(self, four: L.Tensor, full: L.Tensor, direction: int, bias=None, acc: int = 0):
        g = self.geom
        e = self.eng
        L.check(e.lib.semb_pixel_shuffle2x(C.byref(four), C.byref(full), g.N, self.s2d_hw[0], self.s2d_hw[1], g.H, g.W, bias,
                                           direction, acc, e.dtype, e.stream))

    def _fwd_s2d(self, training: bool, sp, ns, cs, bias):
        e = self.eng
        b2 = self.b2.view()
        if not self.transposed:
            self._shuffle(b2.t, self.x.t, 1)                                                   # space-to-depth of the input
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom_s), C.byref(b2.t), self.pk_s["buf"].data_ptr(), bias, C.byref(self.y.t),
                                             sp, ns, cs, 0, e.stream))
            return
        t4, n4, c4 = (None, 0, 0)
        if sp is not None:
            c4 = 4 * self.geom.Cin
            n4 = 2 * c4 if self.stats4_groups > 1 else 0
            t4 = e.zeroed.ptr(self.stats4)
        L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom_t), C.byref(self.x.t), self.pk_t["buf"].data_ptr(), None, C.byref(b2.t),
                                         t4, n4, c4, 0, e.stream))
        self._shuffle(b2.t, self.y.t, 0, bias=bias)                                            # depth-to-space (+ bias)
        if sp is not None:
            L.check(e.lib.semb_fold_stats4(t4, sp, self.stats4_groups, self.geom.Cin, ns, cs, e.stream))

    def _bwd_s2d(self):
        e = self.eng
        dbias = e.gptr(self.bias) if self.bias else None
        b2 = self.b2.view()
        dw3 = self.s2d["dw3"].data_ptr()
        if not self.transposed:
            e.on_wgrad_stream(lambda: e.wgrad_tc(self.geom_s, b2.t, self.y.g, dw3))
            if dbias:
                L.check(e.lib.semb_channel_sum(C.byref(self.y.g), self.geom.N, self.geom.OH * self.geom.OW, dbias, e.dtype, e.stream))
            if self.x.requires_grad:
                L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom_t), C.byref(self.y.g), self.pk_t["buf"].data_ptr(), None, C.byref(b2.g),
                                                 None, 0, 0, 0, e.stream))
                self._shuffle(b2.g, self.x.g, 0, acc=self.acc_x)
            return
        self._shuffle(b2.g, self.y.g, 1)                                                       # space-to-depth of the output gradient
        e.on_wgrad_stream(lambda: e.wgrad_tc(self.geom_s, b2.g, self.x.t, dw3))
        if dbias:
            L.check(e.lib.semb_channel_sum(C.byref(self.y.g), self.geom.N, self.geom.H * self.geom.W, dbias, e.dtype, e.stream))
        if self.x.requires_grad:
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom_s), C.byref(b2.g), self.pk_s["buf"].data_ptr(), None, C.byref(self.x.g),
                                             None, 0, 0, self.acc_x, e.stream))

    def _fwd_tapfold(self, sp, ns, cs, bias):
        e, g = self.eng, self.geom
        mid = self.tf_mid.view()
        if self.tf_valid:
            xp = self.x
        else:
            xp = self.tf_xpad.view()
            L.check(e.lib.semb_pad_crop(C.byref(self.x.t), C.byref(xp.t), g.N, g.H, g.W, g.H + g.R - 1, g.W + g.S - 1, g.pad_t, g.pad_l, 0,
                                        e.dtype, 0, e.stream))
        if self.tapfold["kind"] == "stem":
            L.check(e.lib.semb_tap_patch(C.byref(mid.t), C.byref(xp.t), g.N, g.H, g.W, g.R, None, 0, e.dtype, e.stream))
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom1), C.byref(mid.t), self.pk1["buf"].data_ptr(), bias, C.byref(self.y.t),
                                             sp, ns, cs, 0, e.stream))
        else:
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom1), C.byref(xp.t), self.pk1["buf"].data_ptr(), None, C.byref(mid.t),
                                             None, 0, 0, 0, e.stream))
            L.check(e.lib.semb_tap_patch(C.byref(self.y.t), C.byref(mid.t), g.N, g.OH, g.OW, g.R, bias, 2, e.dtype, e.stream))

    def _bwd_tapfold(self):
        e, g = self.eng, self.geom
        mid = self.tf_mid.view()
        xp = self.x if self.tf_valid else self.tf_xpad.view()
        dw1 = self.tapfold["dw3"].data_ptr()
        dbias = e.gptr(self.bias) if self.bias else None
        if dbias:
            L.check(e.lib.semb_channel_sum(C.byref(self.y.g), g.N, g.OH * g.OW, dbias, e.dtype, e.stream))
        if self.tapfold["kind"] == "stem":
            e.on_wgrad_stream(lambda: e.wgrad_tc(self.geom1, mid.t, self.y.g, dw1))
            if not self.x.requires_grad:
                return
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom1d), C.byref(self.y.g), self.pk1d["buf"].data_ptr(), None, C.byref(mid.g),
                                             None, 0, 0, 0, e.stream))
            L.check(e.lib.semb_tap_patch(C.byref(mid.g), C.byref(xp.g), g.N, g.H, g.W, g.R, None, 1, e.dtype, e.stream))
        else:
            L.check(e.lib.semb_tap_patch(C.byref(self.y.g), C.byref(mid.g), g.N, g.OH, g.OW, g.R, None, 3, e.dtype, e.stream))
            e.on_wgrad_stream(lambda: e.wgrad_tc(self.geom1, xp.t, mid.g, dw1))
            if not self.x.requires_grad:
                return
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom1d), C.byref(mid.g), self.pk1d["buf"].data_ptr(), None, C.byref(xp.g),
                                             None, 0, 0, self.acc_x if self.tf_valid else 0, e.stream))
            if self.tf_valid:
                return          # the data gradient was written (accumulated) straight into x.g: no padding to fold back
        L.check(e.lib.semb_pad_crop(C.byref(xp.g), C.byref(self.x.g), g.N, g.H + g.R - 1, g.W + g.S - 1, g.H, g.W, g.pad_t, g.pad_l, 3,
                                    e.dtype, self.acc_x, e.stream))

    def _bwd_split(self):
        """Parity mode: dW = x^T dy term by term (hh, hm, mh, hl, lh, mm: six tensor-core weight-gradient launches accumulating in
        fp32), dx through the flipped stacked kernel on the split gradient."""
        e, sd, g = self.eng, self.split, self.geom
        cin, cout, T = g.Cin, g.Cout, sd["T"]
        x3 = e.split_scratch(0, T * cin, sd["nx"] * T * cin)
        dy3 = e.split_scratch(1, T * cout, sd["ny"] * T * cout)
        L.check(e.lib.semb_split_bf16(C.byref(self.x.t), C.byref(x3), sd["nx"], e.stream))       # the scratch is shared: split again
        L.check(e.lib.semb_split_bf16(C.byref(self.y.g), C.byref(dy3), sd["ny"], e.stream))
        if not e.skip_wgrad:
            # channel block of the high / middle / low term inside the stacked operand
            hb, mb, lb = (0, None, 1) if T == 3 else (0, 2, 4)
            xs = lambda b: L.Tensor(x3.ptr, cin, T * cin, b * cin)
            ds = lambda b: L.Tensor(dy3.ptr, cout, T * cout, b * cout)
            pairs = [(hb, hb), (lb, hb), (hb, lb)] if T == 3 else [(hb, hb), (hb, mb), (mb, hb), (hb, lb), (lb, hb), (mb, mb)]
            for bx, bd in pairs:
                a_, b_ = xs(bx), ds(bd)
                L.check(e.lib.semb_conv2d_wgrad_tc(C.byref(sd["gw"]), C.byref(a_), C.byref(b_), e.gptr(self.w), e.stream))
        if self.x.requires_grad:
            L.check(e.lib.semb_conv2d_fwd_tc_f32(C.byref(sd["gd"]), C.byref(dy3), sd["pkd"]["buf"].data_ptr(), None, C.byref(self.x.g),
                                                 None, 0, 0, self.acc_x, e.stream))

    def fwd(self, training: bool):
        e = self.eng
        sp, ns, cs = self._stats_args() if training else (None, 0, 0)
        bias = e.params.ptr(self.bias) if self.bias else None
        if self.tapfold is not None:
            return self._fwd_tapfold(sp, ns, cs, bias)
        if self.s2d is not None:
            return self._fwd_s2d(training, sp, ns, cs, bias)
        if self.use_tc and self.x_pad is not None:
            g = self.geom
            xp = self.x_pad.view()
            L.check(e.lib.semb_pad_crop(C.byref(self.x.t), C.byref(xp.t), g.N, g.H, g.W, self.x_pad_hw[0], self.x_pad_hw[1],
                                        g.pad_t, g.pad_l, 0, e.dtype, 0, e.stream))
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom_v), C.byref(xp.t), self.pk_fwd["buf"].data_ptr(), bias,
                                             C.byref(self.y.t), sp, ns, cs, 0, e.stream))
            return
        if self.use_tc and self.d2s_up is not None:
            # Conv2DTranspose(2x2, s2): the conv's epilogue scatters straight into the up-sampled tensor and adds the bias
            up, up_bias, uh, uw = self.d2s_up
            L.check(e.lib.semb_conv2d_fwd_tc_d2s(C.byref(self.geom), C.byref(self.x.t), self.pk_fwd["buf"].data_ptr(),
                                                 e.params.ptr(up_bias) if up_bias else None, C.byref(up.t), uh, uw, e.stream))
            return
        if self.use_tc:
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom), C.byref(self.x.t), self.pk_fwd["buf"].data_ptr(), bias,
                                             C.byref(self.y.t), sp, ns, cs, 0, e.stream))
            return
        if self.split is not None:
            sd = self.split
            T = sd["T"]
            x3 = e.split_scratch(0, T * self.geom.Cin, sd["nx"] * T * self.geom.Cin)
            L.check(e.lib.semb_split_bf16(C.byref(self.x.t), C.byref(x3), sd["nx"], e.stream))
            L.check(e.lib.semb_conv2d_fwd_tc_f32(C.byref(sd["g3"]), C.byref(x3), sd["pk"]["buf"].data_ptr(), None, C.byref(self.y.t),
                                                 sp, ns, cs, 0, e.stream))
            return
        fn = e.lib.semb_conv2d_dgrad if self.transposed else e.lib.semb_conv2d_fwd
        L.check(fn(C.byref(self.geom), C.byref(self.x.t), e.params.ptr(self.w), bias, C.byref(self.y.t), sp, ns, cs, 0,
                   e.stream))

    def bwd(self):
        e = self.eng
        if self.tapfold is not None:
            return self._bwd_tapfold()
        if self.s2d is not None:
            return self._bwd_s2d()
        dbias = e.gptr(self.bias) if self.bias else None
        if self.split is not None:
            return self._bwd_split()
        if not self.transposed:
            if self.use_tc and dbias is None and self.x_pad is not None:
                e.on_wgrad_stream(lambda: e.wgrad_tc(self.geom_v, self.x_pad.view().t, self.y.g, e.gptr(self.w)))
            elif self.use_tc and dbias is None:
                e.on_wgrad_stream(lambda: e.wgrad_tc(self.geom, self.x.t, self.y.g, e.gptr(self.w)))
            else:
                e.on_wgrad_stream(lambda: L.check(e.lib.semb_conv2d_wgrad(C.byref(self.geom), C.byref(self.x.t), C.byref(self.y.g),
                                                                          e.gptr(self.w), dbias, e.stream)))
            if self.x.requires_grad:
                dst = self.x.g if self.pad_buf is None else self.pad_buf.view().t
                acc = self.acc_x if self.pad_buf is None else 0
                if self.use_tc and self.pk_bwd is not None:
                    L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom_d), C.byref(self.y.g), self.pk_bwd["buf"].data_ptr(), None,
                                                     C.byref(dst), None, 0, 0, acc, e.stream))
                else:
                    L.check(e.lib.semb_conv2d_dgrad(C.byref(self.geom if self.pad_buf is None else self.geom_p), C.byref(self.y.g),
                                                    e.params.ptr(self.w), None, C.byref(dst), None, 0, 0, acc, e.stream))
                if self.pad_buf is not None:
                    g = self.geom
                    L.check(e.lib.semb_pad_crop(C.byref(dst), C.byref(self.x.g), g.N, self.pad_hw[0], self.pad_hw[1], g.H, g.W,
                                                g.pad_t, g.pad_l, 3, e.dtype, self.acc_x, e.stream))
        else:
            # d/dw of the transposed conv: wgrad of the equivalent conv with x':=d(out), dy':=in
            L.check(e.lib.semb_conv2d_wgrad(C.byref(self.geom), C.byref(self.y.g), C.byref(self.x.t), e.gptr(self.w), None,
                                            e.stream))
            if dbias:
                L.check(e.lib.semb_channel_sum(C.byref(self.y.g), self.geom.N, self.geom.H * self.geom.W, dbias, e.dtype,
                                               e.stream))
            if self.x.requires_grad:
                L.check(e.lib.semb_conv2d_fwd(C.byref(self.geom), C.byref(self.y.g), e.params.ptr(self.w), None,
                                              C.byref(self.x.g), None, 0, 0, self.acc_x, e.stream))


class PairConvOp(ConvOp):
    """res_path unit (UNet_Segmentation.py:490-499): `Conv2D(3x3)` and the 1x1 shortcut `Conv2D` of the SAME input as ONE 3x3
    tensor-core conv with Ca + Cs output channels (the shortcut kernel sits on the centre tap of the virtual kernel, see
    semb_merge_weights).  y holds [3x3 output | shortcut output]; its gradient is consumed by ONE data-gradient launch (no
    read-modify-write of dx) and ONE weight-gradient launch whose result is folded into the two Keras-layout gradients.
    bf16 tensor-core mode only; the builders fall back to two ConvOps elsewhere."""

    def __init__(self, eng: Engine, x: View, y: View, hw: Tuple[int, int], w_a: str, w_s: str, ca: int, cs: int,
                 stats: Optional[Tuple[str, int, int, int]] = None):
        assert eng.tc_enabled and y.C == ca + cs
        self.eng, self.x, self.y, self.w, self.w_s, self.bias, self.transposed = eng, x, y, w_a, w_s, None, False
        h, wd = hw
        n, cin, cout = eng.N, x.C, ca + cs
        self.geom = L.ConvGeom(n, h, wd, h, wd, cin, cout, 3, 3, 1, 1, 1, L.PAD_ZERO, eng.dtype)
        self.stats = stats
        self.acc_x = 0
        self.use_tc = True
        self.pad_buf = self.s2d = self.tapfold = self.x_pad = self.split = None
        self.rec = eng.merge_weight(w_a, w_s, cin, ca, cs)
        self.pk_fwd = eng.tc_pack(self.rec["key"], 3, 3, cin, cout, 0, vw=self.rec)
        self.pk_bwd = None
        if x.requires_grad:
            self.pk_bwd = eng.tc_pack(self.rec["key"], 3, 3, cin, cout, 1, vw=self.rec)
            self.geom_d = L.ConvGeom(n, h, wd, h, wd, cout, cin, 3, 3, 1, 1, 1, L.PAD_ZERO, eng.dtype)
        # algorithmic work of the two reference layers (the zero taps of the shortcut columns are not counted)
        self.alg_flops = 2.0 * n * h * wd * cin * (9 * ca + cs)

    def fwd(self, training: bool):
        e = self.eng
        sp, ns, cs = self._stats_args() if training else (None, 0, 0)
        L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom), C.byref(self.x.t), self.pk_fwd["buf"].data_ptr(), None, C.byref(self.y.t),
                                         sp, ns, cs, 0, e.stream))

    def bwd(self):
        e = self.eng
        e.on_wgrad_stream(lambda: e.wgrad_tc(self.geom, self.x.t, self.y.g, self.rec["dw3"].data_ptr()))
        if self.x.requires_grad:
            L.check(e.lib.semb_conv2d_fwd_tc(C.byref(self.geom_d), C.byref(self.y.g), self.pk_bwd["buf"].data_ptr(), None, C.byref(self.x.g),
                                             None, 0, 0, self.acc_x, e.stream))


class NormOp(Op):
    """Turns accumulated moments into (scale, shift, mean, invstd); BatchNorm (groups=1) or InstanceNorm (groups=N).

    Owns the per-layer scratch.  Has no backward of its own: the gradient w.r.t. gamma/beta and the
    normalisation terms are produced by the AffineOp that applies this norm."""

    def __init__(self, eng: Engine, name: str, c: int, count: float, eps: float, gamma: Optional[str], beta: str,
                 moving: Optional[Tuple[str, str]], momentum: float, groups: int = 1):
        self.eng, self.name, self.C, self.count, self.eps = eng, name, c, float(count), eps
        self.gamma, self.beta, self.moving, self.momentum, self.groups = gamma, beta, moving, momentum, groups
        self.stats = eng.zeroed.add(name + "/moments", 4 * c * groups)   # fp64 accumulators: 2 moments x C x groups
        self.sums = eng.zeroed.add(name + "/bwd_sums", 4 * c * groups)
        for k in ("scale", "shift", "mean", "invstd", "c1", "c2"):
            eng.scratch.add(f"{name}/{k}", c * groups)
        self.nstride = 0 if groups == 1 else c
        self.stats_nstride = 0 if groups == 1 else 2 * c
        self.consumers = 0       # AffineOps applying this norm; with exactly one the finalize is folded into its kernel
        # where the fp64 moments are: (zeroed-store name, channel offset, stride between the sum and the sum of squares);
        # a PairConvOp writes the moments of two norms into ONE block (bind_stats)
        self.stats_off, self.stats_cs = 0, c

    def bind_stats(self, name: str, off: int, cstride: int):
        """Moments come from a block shared with another norm: channel `off` of a `cstride`-channel record (BatchNorm only)."""
        assert self.groups == 1
        self.stats, self.stats_off, self.stats_cs = name, off, cstride

    def s(self, k: str) -> int:
        return self.eng.scratch.ptr(f"{self.name}/{k}")

    def stats_ref(self, coff: int = 0):
        return (self.stats, self.stats_off + coff, self.stats_nstride, self.stats_cs)

    def folded(self, training: bool) -> bool:
        return self.eng.fold_norm and self.consumers == 1 and self.uses_batch_stats(training)

    def fin(self, training: bool, coff: int = 0) -> L.NormFin:
        """semb_norm_fin record for the consumer kernel (stats=NULL when the finalize is not folded)."""
        e = self.eng
        f = L.NormFin()
        f.scale, f.shift = self.s("scale") + 4 * coff, self.s("shift") + 4 * coff
        if self.folded(training):
            f.stats = e.zeroed.ptr(self.stats, 2 * (self.stats_off + coff))
            f.stats_nstride, f.cstride = self.stats_nstride, self.stats_cs
            f.count, f.eps = self.count, self.eps
            f.gamma = (e.params.ptr(self.gamma) + 4 * coff) if self.gamma else None
            f.beta = e.params.ptr(self.beta) + 4 * coff
            f.mean, f.invstd = self.s("mean") + 4 * coff, self.s("invstd") + 4 * coff
            if self.moving and training:
                f.moving_mean, f.moving_var = e.state.ptr(self.moving[0]) + 4 * coff, e.state.ptr(self.moving[1]) + 4 * coff
            f.momentum = self.momentum
        return f

    def fwd(self, training: bool):
        e = self.eng
        if self.folded(training):
            return          # the AffineOp that applies this norm finalizes it in its own kernel
        gamma = e.params.ptr(self.gamma) if self.gamma else None
        beta = e.params.ptr(self.beta)
        if training or self.moving is None:
            mm = e.state.ptr(self.moving[0]) if (self.moving and training) else None
            mv = e.state.ptr(self.moving[1]) if (self.moving and training) else None
            L.check(e.lib.semb_norm_finalize(e.zeroed.ptr(self.stats, 2 * self.stats_off), self.groups, self.C, self.stats_cs, self.stats_nstride,
                                             self.count, self.eps, gamma, beta, self.s("scale"), self.s("shift"),
                                             self.s("mean"), self.s("invstd"), mm, mv, self.momentum, e.stream))
        else:
            L.check(e.lib.semb_norm_from_moving(self.C, self.eps, gamma, beta, e.state.ptr(self.moving[0]),
                                                e.state.ptr(self.moving[1]), self.s("scale"), self.s("shift"), e.stream))

    def uses_batch_stats(self, training: bool) -> bool:
        return training or self.moving is None


class SegView:
    """The b operand of an AffineOp as a channel-wise concatenation of compact Views (no concat buffer): quacks like a
    View where the kernels only need the total channel count (semb_affine_desc.nseg_b)."""

    def __init__(self, segs: Sequence[View]):
        assert 1 <= len(segs) <= 3
        self.segs = list(segs)
        self.C = sum(v.C for v in segs)
        self.c0 = [sum(v.C for v in segs[:i]) for i in range(len(segs))]
        self.requires_grad = any(v.requires_grad for v in segs)
        self._t = self._g = None

    @property
    def t(self) -> L.Tensor:            # placeholder (valid view over the first segment's storage, never dereferenced)
        if self._t is None:
            self._t = L.Tensor(self.segs[0].buf.data.data_ptr(), self.C, self.C, 0)
        return self._t

    @property
    def g(self) -> L.Tensor:
        if self._g is None:
            self._g = L.Tensor(self.segs[0].buf.grad_tensor().data_ptr(), self.C, self.C, 0)
        return self._g


class AffineOp(Op):
    """y = act( A(a) + actb(B(b)) ) with A/B the affines of NormOps (or identity), optional moments of y.
    `b` may be a SegView: the concatenation of up to three compact tensors (multi_res_block's `concatenate`)."""

    def __init__(self, eng: Engine, hw: int, a: View, norm_a: Optional[NormOp], b: Optional[View], norm_b: Optional[NormOp],
                 y: View, act: int, actb: int = L.ACT_NONE, stats_out: Optional[Tuple[str, int, int, int]] = None,
                 n: Optional[int] = None):
        self.eng, self.a, self.b, self.y, self.norm_a, self.norm_b = eng, a, b, y, norm_a, norm_b
        self.act, self.actb, self.stats_out = act, actb, stats_out
        self.coff_a = 0
        self.hw = hw
        self.n = eng.N if n is None else n
        self.acc_a = self.acc_b = 0
        groups = max(norm_a.groups if norm_a else 1, norm_b.groups if norm_b else 1)
        self.aff_nstride = 0 if groups == 1 else (norm_a or norm_b).C
        # two zeroed words for the grid barrier of the fused backward kernel
        for nm in (norm_a, norm_b):
            if nm is not None:
                nm.consumers += 1
        eng._naff = getattr(eng, "_naff", 0) + 1
        self.bar = eng.zeroed.add(f"affine_{eng._naff}/barrier", 4)

    def plan_backward(self):
        if self.a.requires_grad:
            self.acc_a = plan_grad_write(self.a)
        if isinstance(self.b, SegView):
            accs = {plan_grad_write(v) for v in self.b.segs if v.requires_grad}
            assert len(accs) <= 1, "segments of one concatenation must all be first (or all later) gradient writers"
            self.acc_b = accs.pop() if accs else 0
        elif self.b is not None and self.b.requires_grad:
            self.acc_b = plan_grad_write(self.b)

    def _desc(self, training: bool) -> L.AffineDesc:
        cache = self.__dict__.setdefault("_desc_cache", {})
        d = cache.get(training)
        if d is None:
            d = cache[training] = self._make_desc(training)
        return d

    def _make_desc(self, training: bool) -> L.AffineDesc:
        def mode(norm):
            if norm is None:
                return L.AFF_NONE
            return L.AFF_BATCH if norm.uses_batch_stats(training) else L.AFF_PLAIN
        d = L.AffineDesc(self.n, self.hw, self.a.C, self.eng.dtype, self.act, self.actb, mode(self.norm_a),
                         mode(self.norm_b) if self.b is not None else L.AFF_NONE, self.aff_nstride)
        if isinstance(self.b, SegView):
            d.nseg_b = len(self.b.segs)
            for i, v in enumerate(self.b.segs):
                d.seg_c0[i] = self.b.c0[i]
                d.seg_b[i] = v.t
        return d

    def _p(self, norm: Optional[NormOp], key: str, coff: int) -> Optional[int]:
        return None if norm is None else norm.s(key) + 4 * coff

    def fwd(self, training: bool):
        e = self.eng
        d = self._desc(training)
        self._last_desc = d
        sp, ns, cs = (None, 0, 0)
        if self.stats_out is not None and training:
            name, off, ns, cs = self.stats_out
            sp = e.zeroed.ptr(name, 2 * off)
        if e.fold_norm:
            fa = self.norm_a.fin(training, self.coff_a) if self.norm_a is not None else None
            fb = self.norm_b.fin(training, 0) if (self.norm_b is not None and self.b is not None) else None
            L.check(e.lib.semb_affine_act_fwd_fin(
                C.byref(d), C.byref(self.a.t), C.byref(fa) if fa is not None else None,
                C.byref(self.b.t) if self.b is not None else None, C.byref(fb) if fb is not None else None,
                C.byref(self.y.t), sp, ns, cs, e.stream))
            return
        L.check(e.lib.semb_affine_act_fwd(
            C.byref(d), C.byref(self.a.t), self._p(self.norm_a, "scale", self.coff_a), self._p(self.norm_a, "shift", self.coff_a),
            C.byref(self.b.t) if self.b is not None else None, self._p(self.norm_b, "scale", 0), self._p(self.norm_b, "shift", 0),
            C.byref(self.y.t), sp, ns, cs, e.stream))

    def bwd(self):
        e = self.eng
        d = self._last_desc
        if isinstance(self.b, SegView):
            for i, v in enumerate(self.b.segs):
                if v.requires_grad:
                    d.seg_db[i] = v.g
        na, nb = self.norm_a, self.norm_b
        ca = self.coff_a
        bt = C.byref(self.b.t) if self.b is not None else None
        red_a = d.mode_a == L.AFF_BATCH
        red_b = self.b is not None and d.mode_b == L.AFF_BATCH
        # the bwd sums of this op live in norm_a's (resp. norm_b's) scratch; slot [0,1] = a-terms, [2,3] = b-terms
        da = C.byref(self.a.g) if self.a.requires_grad else None
        dbv = C.byref(self.b.g) if (self.b is not None and self.b.requires_grad) else None
        if (red_a or red_b) and e.fused_affine_bwd:
            owner = na if red_a else nb
            sums = e.zeroed.ptr(owner.sums, ca if red_a else 0)
            cs = owner.C
            ns = 0 if owner.groups == 1 else 4 * owner.C
            dga = (e.gptr(na.gamma) + 4 * ca) if (red_a and na.gamma) else None
            dba = (e.gptr(na.beta) + 4 * ca) if red_a else None
            dgb = e.gptr(nb.gamma) if (red_b and nb.gamma) else None
            dbb = e.gptr(nb.beta) if red_b else None
            rc = e.lib.semb_affine_act_bwd_fused(
                C.byref(d), C.byref(self.y.g), C.byref(self.a.t), bt,
                self._p(na, "scale", ca), self._p(na, "shift", ca), self._p(na, "mean", ca), self._p(na, "invstd", ca),
                na.count if na else 0.0, dga, dba,
                self._p(nb, "scale", 0), self._p(nb, "shift", 0), self._p(nb, "mean", 0), self._p(nb, "invstd", 0),
                nb.count if nb else 0.0, dgb, dbb,
                sums, ns, cs, e.zeroed.ptr(self.bar), da, self.acc_a, dbv, self.acc_b, e.stream)
            if rc == 0:
                return
            if rc != -4:            # SEMB_EWORKSPACE: grid not co-resident -> two-pass kernels below
                L.check(rc)
        if red_a or red_b:
            owner = na if red_a else nb
            sums = e.zeroed.ptr(owner.sums, ca if red_a else 0)
            cs = owner.C
            ns = 0 if owner.groups == 1 else 4 * owner.C
            L.check(e.lib.semb_affine_act_bwd_reduce(
                C.byref(d), C.byref(self.y.g), C.byref(self.a.t), bt,
                self._p(na, "scale", ca), self._p(na, "shift", ca), self._p(na, "mean", ca), self._p(na, "invstd", ca),
                self._p(nb, "scale", 0), self._p(nb, "shift", 0), self._p(nb, "mean", 0), self._p(nb, "invstd", 0),
                sums, ns, cs, e.stream))
            if e.fold_norm:
                # pass 2 reads the sums directly and accumulates dgamma / dbeta (semb_norm_bwd_finalize folded in)
                dga = (e.gptr(na.gamma) + 4 * ca) if (red_a and na.gamma) else None
                dba = (e.gptr(na.beta) + 4 * ca) if red_a else None
                dgb = e.gptr(nb.gamma) if (red_b and nb.gamma) else None
                dbb = e.gptr(nb.beta) if red_b else None
                L.check(e.lib.semb_affine_act_bwd_apply_sums(
                    C.byref(d), C.byref(self.y.g), C.byref(self.a.t), bt,
                    self._p(na, "scale", ca), self._p(na, "shift", ca), self._p(na, "mean", ca), self._p(na, "invstd", ca),
                    na.count if na else 0.0, dga, dba,
                    self._p(nb, "scale", 0), self._p(nb, "shift", 0), self._p(nb, "mean", 0), self._p(nb, "invstd", 0),
                    nb.count if nb else 0.0, dgb, dbb, sums, ns, cs, da, self.acc_a, dbv, self.acc_b, e.stream))
                return
            if red_a:
                dg = e.gptr(na.gamma) + 4 * ca if na.gamma else None
                db = e.gptr(na.beta) + 4 * ca
                L.check(e.lib.semb_norm_bwd_finalize(sums, 0, na.groups, self.a.C, cs, ns, na.count, None, None,
                                                     self._p(na, "c1", ca), self._p(na, "c2", ca), dg, db, e.stream))
            if red_b:
                dg = e.gptr(nb.gamma) if nb.gamma else None
                db = e.gptr(nb.beta)
                # b-terms were written at slots 2,3 of `sums`; c1/c2 of norm_b use its own arrays.  When the owner of
                # the sums buffer is norm_a (different C stride) the b-sums still sit at 2*cs, 3*cs.
                L.check(e.lib.semb_norm_bwd_finalize(sums, 1, nb.groups, self.a.C, cs, ns, nb.count, None, None,
                                                     self._p(nb, "c1", 0), self._p(nb, "c2", 0), dg, db, e.stream))
        if da is None and dbv is None:
            return
        L.check(e.lib.semb_affine_act_bwd_apply(
            C.byref(d), C.byref(self.y.g), C.byref(self.a.t), bt,
            self._p(na, "scale", ca), self._p(na, "shift", ca), self._p(na, "mean", ca), self._p(na, "invstd", ca),
            self._p(na, "c1", ca), self._p(na, "c2", ca),
            self._p(nb, "scale", 0), self._p(nb, "shift", 0), self._p(nb, "mean", 0), self._p(nb, "invstd", 0),
            self._p(nb, "c1", 0), self._p(nb, "c2", 0), da, self.acc_a, dbv, self.acc_b, e.stream))


class PoolOp(Op):
    def __init__(self, eng: Engine, x: View, y: View, h: int, w: int):
        self.eng, self.x, self.y, self.h, self.w = eng, x, y, h, w
        self.acc = 0

    def plan_backward(self):
        self.acc = plan_grad_write(self.x)

    def fwd(self, training: bool):
        e = self.eng
        L.check(e.lib.semb_maxpool2x2_fwd(C.byref(self.x.t), C.byref(self.y.t), e.N, self.h, self.w, e.dtype, e.stream))

    def bwd(self):
        e = self.eng
        L.check(e.lib.semb_maxpool2x2_bwd(C.byref(self.x.t), C.byref(self.y.g), C.byref(self.x.g), e.N, self.h, self.w,
                                          e.dtype, self.acc, e.stream))


class PadCropOp(Op):
    """mode 'reflect' (ReflectionPadding2D) or 'crop' (Cropping2D)."""

    def __init__(self, eng: Engine, x: View, y: View, hw_in, hw_out, top: int, left: int, mode: str):
        self.eng, self.x, self.y, self.hw_in, self.hw_out, self.top, self.left, self.mode = eng, x, y, hw_in, hw_out, top, left, mode
        self.acc = 0

    def plan_backward(self):
        if self.x.requires_grad:
            self.acc = plan_grad_write(self.x)

    def fwd(self, training: bool):
        e = self.eng
        m = 0 if self.mode == "reflect" else 1
        L.check(e.lib.semb_pad_crop(C.byref(self.x.t), C.byref(self.y.t), e.N, self.hw_in[0], self.hw_in[1], self.hw_out[0],
                                    self.hw_out[1], self.top, self.left, m, e.dtype, 0, e.stream))

    def bwd(self):
        if not self.x.requires_grad:
            return
        e = self.eng
        m = 3 if self.mode == "reflect" else 2
        L.check(e.lib.semb_pad_crop(C.byref(self.y.g), C.byref(self.x.g), e.N, self.hw_out[0], self.hw_out[1], self.hw_in[0],
                                    self.hw_in[1], self.top, self.left, m, e.dtype, self.acc, e.stream))


class ShuffleOp(Op):
    """Depth-to-space half of Conv2DTranspose(2x2, stride 2): y4 (N,H,W,4C) -> out (N,2H,2W,C) + bias."""

    def __init__(self, eng: Engine, y4: View, out: View, h: int, w: int, bias: Optional[str], conv: Optional["ConvOp"] = None):
        self.eng, self.y4, self.out, self.h, self.w, self.bias = eng, y4, out, h, w, bias
        # bf16 tensor-core mode: the producing 1x1 conv scatters into `out` itself (semb_conv2d_fwd_tc_d2s); this op then
        # only exists for the backward pass (gather of d(out) into d(y4), bias gradient)
        self.fused_fwd = (conv is not None and conv.use_tc and conv.stats is None and conv.bias is None and conv.x_pad is None
                          and conv.split is None and conv.s2d is None and conv.tapfold is None and conv.geom.R == 1
                          and _os.environ.get("SEMB_NO_D2S_FUSE") is None)
        if self.fused_fwd:
            conv.d2s_up = (out, bias, 2 * h, 2 * w)

    def plan_backward(self):
        acc = plan_grad_write(self.y4)
        assert acc == 0

    def fwd(self, training: bool):
        e = self.eng
        if self.fused_fwd:
            return
        L.check(e.lib.semb_pixel_shuffle2(C.byref(self.y4.t), C.byref(self.out.t), e.N, self.h, self.w,
                                          e.params.ptr(self.bias) if self.bias else None, 0, e.dtype, e.stream))

    def bwd(self):
        e = self.eng
        L.check(e.lib.semb_pixel_shuffle2(C.byref(self.y4.g), C.byref(self.out.g), e.N, self.h, self.w, None, 1, e.dtype, e.stream))
        if self.bias:
            L.check(e.lib.semb_channel_sum(C.byref(self.out.g), e.N, 4 * self.h * self.w, e.gptr(self.bias), e.dtype, e.stream))


class UpsampleOp(Op):
    """UpSampling2D(size=(2,2)), nearest neighbour (CycleGAN.py:349)."""

    def __init__(self, eng: Engine, x: View, y: View, h: int, w: int, n: Optional[int] = None):
        self.eng, self.x, self.y, self.h, self.w = eng, x, y, h, w
        self.n = eng.N if n is None else n
        self.acc = 0

    def plan_backward(self):
        if self.x.requires_grad:
            self.acc = plan_grad_write(self.x)

    def fwd(self, training: bool):
        e = self.eng
        L.check(e.lib.semb_upsample2x(C.byref(self.x.t), C.byref(self.y.t), self.n, self.h, self.w, 0, 0, e.dtype, e.stream))

    def bwd(self):
        if not self.x.requires_grad:
            return
        e = self.eng
        L.check(e.lib.semb_upsample2x(C.byref(self.x.g), C.byref(self.y.g), self.n, self.h, self.w, 1, self.acc, e.dtype, e.stream))


class MaskOp(Op):
    """y = x * slope(z) * m: LeakyReLU and/or inverted Dropout as ONE multiplicative mask (semb_mask_mul; WassersteinGAN.py
    conv_block :547-567, Dropout :618).  z = the pre-activation whose sign selects slope 1 / `neg` (None: no activation),
    m = the keep mask in `drop` (values 0 or 1/(1-rate); None: no dropout).  `like` = another MaskOp whose z and m this op
    re-uses: the critic linearised at that tower's operating point (the gradient-penalty tower).

    The keep mask is drawn by torch's device generator (random bits, not arithmetic of the path); `frozen=True` keeps the
    current content of `drop` (parity tests feed the oracle the same masks); training=False makes Dropout the identity."""

    def __init__(self, eng: Engine, x: View, y: View, npix: int, z: Optional[View] = None, rate: float = 0.0, neg: float = 0.2,
                 like: Optional["MaskOp"] = None, n: Optional[int] = None):
        self.eng, self.x, self.y, self.npix, self.neg, self.like = eng, x, y, npix, float(neg), like
        self.z = like.z if like is not None else z
        self.rate = like.rate if like is not None else float(rate)
        self.drop = None
        if like is not None:
            self.drop = like.drop
        elif self.rate > 0:
            self.drop = eng.new_buf(x.buf.H, x.buf.W, x.C, f"dropmask_{len(eng.ops)}", requires_grad=False, n=x.buf.N)
        self.frozen = False
        self.acc_x = 0
        self._training = True

    def plan_backward(self):
        if self.x.requires_grad:
            self.acc_x = plan_grad_write(self.x)

    def _m(self):
        return C.byref(self.drop.view().t) if (self.drop is not None and self._training) else None

    def fwd(self, training: bool):
        e = self.eng
        self._training = training
        if self.like is None and self.drop is not None and training and not self.frozen:
            d = self.drop.data
            d.copy_((torch.rand(d.shape, device=d.device) >= self.rate).to(d.dtype) * (1.0 / (1.0 - self.rate)))
        elif self.like is not None:
            self._training = self.like._training
        L.check(e.lib.semb_mask_mul(C.byref(self.x.t), C.byref(self.z.t) if self.z is not None else None, self._m(), C.byref(self.y.t),
                                    self.npix, self.neg, 0, e.dtype, e.stream))

    def bwd(self):
        e = self.eng
        if not self.x.requires_grad:
            return
        L.check(e.lib.semb_mask_mul(C.byref(self.y.g), C.byref(self.z.t) if self.z is not None else None, self._m(), C.byref(self.x.g),
                                    self.npix, self.neg, self.acc_x, e.dtype, e.stream))


class NoiseOp(Op):
    """keras.layers.GaussianNoise(stddev) (CycleGAN.py:427-447): y = x + N(0, stddev) while training, identity otherwise.

    The normal deviates are drawn by torch's device generator into `noise` (plumbing: random bits, not arithmetic of the
    path); the add and its gradient run in the affine kernels.  `frozen=True` keeps the current content of `noise`
    (parity tests feed the same deviates to the oracle)."""

    def __init__(self, eng: Engine, x: View, y: View, hw: int, stddev: float, n: Optional[int] = None, c_logical: Optional[int] = None):
        self.eng, self.x, self.y, self.stddev = eng, x, y, float(stddev)
        self.c_logical = x.C if c_logical is None else c_logical       # padded channel lanes must stay exactly zero
        self.noise = eng.new_buf(x.buf.H, x.buf.W, x.C, f"noise_{len(eng.ops)}", requires_grad=False, n=n)
        self.frozen = False
        self.add = AffineOp(eng, hw, x, None, self.noise.view(), None, y, L.ACT_NONE, n=n)

    def plan_backward(self):
        self.add.plan_backward()

    def fwd(self, training: bool):
        if training and not self.frozen:
            self.noise.data.normal_(0.0, self.stddev)
            if self.c_logical < self.x.C:
                self.noise.data[..., self.c_logical:].zero_()
        elif not training:
            self.noise.data.zero_()
        self.add.fwd(training)

    def bwd(self):
        self.add.bwd()
