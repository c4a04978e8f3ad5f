"""Keras-`Model`-shaped front end over the engine: `model(x, training)`, `get_weights/set_weights`,
`compile`, `train_step`, `fit`, `save` / `load_model`.

It mirrors the protocol the reference relies on (SURVEY.md 8b "Model call protocol"): NHWC float32 in and
out, weights as a list of numpy arrays in keras.Model.layers order, Keras' TorchTrainer.train_step
sequence (forward(training=True) -> loss -> backward -> Adam -> metrics) [K3.5].
"""
from __future__ import annotations

import collections
import collections.abc
import ctypes as C
import json
import math
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .engine import Engine
from .nets import UNetBuilder


def _to_numpy(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class _Instance:
    """One engine specialised to (N, H, W)."""

    def __init__(self, n, h, w, filters, dtype, use_tc=True):
        self.eng = Engine(n, dtype, use_tc=use_tc)
        self.net = UNetBuilder(self.eng, h, w, filters)
        e = self.eng
        self.loss_sums = e.zeroed.add("loss/sums", 4)
        e.finalize()
        if os.environ.get("SEMB_NO_WGRAD_STREAM") is None and not e.tc_split:      # split-operand scratch is shared: one stream
            e.wgrad_stream = torch.cuda.Stream(device=e.device)
            e.enable_lanes()        # res_paths on their own streams (no-op under SEMB_NO_LANES)
        self.n, self.h, self.w = n, h, w
        self.x_dev = torch.zeros((n, h, w, 1), dtype=torch.float32, device=e.device)
        self.y_dev = torch.zeros((n, h, w, 1), dtype=torch.float32, device=e.device)
        self.out_dev = torch.zeros((n, h, w, 1), dtype=torch.float32, device=e.device)
        self.x_pin = torch.zeros((n, h, w, 1), dtype=torch.float32).pin_memory()
        self.out_pin = torch.zeros((n, h, w, 1), dtype=torch.float32).pin_memory()
        self.sums_pin = torch.zeros(4, dtype=torch.float32).pin_memory()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph2: Optional[torch.cuda.CUDAGraph] = None
        self.graph_key = None
        self._train_io = None

    def train_io(self) -> "_TrainIO":
        if self._train_io is None:
            self._train_io = _TrainIO(self)
        return self._train_io

    # ---- pieces of a step (all asynchronous on the current stream) ---------------------------------
    def stage_in(self):
        e = self.eng
        L.check(e.lib.semb_cast_in(self.x_dev.data_ptr(), 1, C.byref(self.net.in_buf.view().t), self.n * self.h * self.w,
                                   e.dtype, e.stream))

    def stage_out(self):
        e = self.eng
        L.check(e.lib.semb_cast_out(C.byref(self.net.out_buf.view().t), self.out_dev.data_ptr(), 1, self.n * self.h * self.w,
                                    e.dtype, e.stream))

    def loss(self, weighting: float, with_grad: bool):
        e = self.eng
        ov = self.net.out_buf.view()
        # p = sigmoid(BN(z)) is recomputed in fp32 from the stored pre-activation z: a bf16 probability saturates at
        # p ~ 0.998, long before Keras' clip at 1 - 1e-7 (UNet_Segmentation.py:379-384)
        hn = self.net.head_norm
        L.check(e.lib.semb_loss_wbce_logits(C.byref(self.net.head_raw.view.t), hn.s("scale"), hn.s("shift"), self.y_dev.data_ptr(),
                                            C.byref(ov.g) if with_grad else None, self.n * self.h * self.w, float(weighting),
                                            e.zeroed.ptr(self.loss_sums), e.dtype, e.stream))

    def fwd_bwd(self, weighting: float):
        """zero scratch -> forward -> loss -> backward.  The input cast (stage_in) is NOT part of it: it runs in front of
        the captured graph so that the host->device copy of the next batch can overlap this step (see _TrainIO)."""
        e = self.eng
        e.zero_step(zero_grads=True)
        e.forward(training=True)
        self.loss(weighting, with_grad=True)
        e.backward()
        if e.s2d:
            e.fold_virtual_grads()      # `grads` must be complete BEFORE the data-parallel all-reduce (Engine.adam folds too late)


class _TrainIO:
    """Double-buffered input staging of one training instance.

    Step i+1's batch travels host -> device on a copy stream while step i's graph runs: pinned (or caller-pinned,
    zero-copy) host memory -> x_stage[slot] / y_stage[slot]; on the compute stream a cast kernel and one device copy move
    the slot into the buffers the graph reads, which frees the slot again long before its next use (two steps later)."""

    RING = 4

    def __init__(self, inst: "_Instance"):
        n, h, w, dev = inst.n, inst.h, inst.w, inst.eng.device
        mk = lambda: torch.zeros((n, h, w, 1), dtype=torch.float32, device=dev)
        self.x_stage, self.y_stage = [mk(), mk()], [mk(), mk()]
        self.x_pin, self.y_pin = [None, None], [None, None]
        self.copy_stream = torch.cuda.Stream(device=dev)
        # The staging buffers above were allocated and zero-filled on the CURRENT stream.  The first upload must not overtake
        # those fills: on a busy main stream the H2D copy on the copy stream used to land first and was then zeroed -- the
        # first train_step of an instance intermittently saw an all-zero batch (found by tests/test_dp_gpu.py on a GPU shared
        # by two processes; every later step was ordered by ev_free).
        self.copy_stream.wait_stream(torch.cuda.current_stream(dev))
        self.ev_h2d = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        self.sums_pin = [torch.zeros(4, dtype=torch.float32).pin_memory() for _ in range(self.RING)]
        self.ev_sums = [torch.cuda.Event() for _ in range(self.RING)]
        self.step = 0
        self.shape = (n, h, w, 1)

    def _host(self, arr, pins, slot) -> torch.Tensor:
        """A pinned float32 view of `arr`: the caller's own tensor when it is already pinned (zero copy), else a memcpy
        into this slot's pinned buffer (after the H2D that last read it has finished)."""
        if isinstance(arr, torch.Tensor) and arr.device.type == "cpu" and arr.dtype == torch.float32 and arr.is_contiguous() \
                and arr.is_pinned() and tuple(arr.shape) == self.shape:
            return arr
        if pins[slot] is None:
            pins[slot] = torch.zeros(self.shape, dtype=torch.float32).pin_memory()
        self.ev_h2d[slot].synchronize()
        src = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(_to_numpy(arr), dtype=np.float32))
        pins[slot].copy_(src.reshape(self.shape))
        return pins[slot]

    def upload(self, x, y) -> int:
        slot = self.step & 1
        xs, ys = self._host(x, self.x_pin, slot), self._host(y, self.y_pin, slot)
        cs = self.copy_stream
        cs.wait_event(self.ev_free[slot])
        with torch.cuda.stream(cs):
            self.x_stage[slot].copy_(xs, non_blocking=True)
            self.y_stage[slot].copy_(ys, non_blocking=True)
            self.ev_h2d[slot].record(cs)
        return slot


class LazyLogs(collections.abc.Mapping):
    """Metrics of one step.  The device->host copy of the loss sums is enqueued with the step; the values are only waited
    for when they are read (Keras' torch trainer also hands back device tensors and converts them when logging)."""

    def __init__(self, pin: torch.Tensor, event: torch.cuda.Event, count: float):
        self._pin, self._event, self._count, self._vals = pin, event, count, None

    def _resolve(self):
        if self._vals is None:
            self._event.synchronize()
            s = self._pin
            self._vals = {"loss": float(s[0]) / self._count, "mae": float(s[1]) / self._count, "acc": float(s[2]) / self._count}
        return self._vals

    def __getitem__(self, k):
        return self._resolve()[k]

    def __iter__(self):
        return iter(("loss", "mae", "acc"))

    def __len__(self):
        return 3

    def __repr__(self):
        return repr(self._resolve())


class UNetModel:
    """MultiRes-UNet with the Keras model protocol, running on libsemb200 (sm_100a) only."""

    def __init__(self, input_shape=(256, 256, 1), filters: int = 16, output_channels: int = 1, dtype: str = "f32",
                 batch_size: int = 1, seed: int = 0, use_cuda_graph: bool = True, use_tc: bool = True):
        if len(input_shape) == 2:
            input_shape = (input_shape[0], input_shape[1], 1)
        assert input_shape[2] == 1 and output_channels == 1
        self.input_shape = tuple(input_shape)
        self.filters, self.dtype, self.seed = filters, dtype, seed
        self.use_cuda_graph = use_cuda_graph
        self.use_tc = use_tc
        # engines are specialised to (N, H, W); keep the training instance plus a few most-recently-used others (a
        # directory of differently sized images must not grow device / pinned memory without bound)
        self._instances: "collections.OrderedDict[tuple, _Instance]" = collections.OrderedDict()
        self.max_instances = 3
        self._primary = None
        self._current = None
        self._primary = self._instance(batch_size, input_shape[0], input_shape[1], init=True)
        self._current = self._primary
        self.weighting = 1.0
        self.learning_rate = 1e-3
        self.beta_1, self.beta_2, self.epsilon = 0.9, 0.999, 1e-7
        self.world_size, self.rank = 1, 0
        self.process_group = None
        self.stop_training = False
        self.history: List[dict] = []

    # ---- instances -----------------------------------------------------------------------------------
    def _instance(self, n, h, w, init=False) -> _Instance:
        key = (n, h, w)
        inst = self._instances.get(key)
        if inst is None:
            # evict least-recently-used instances first (never the training instance or the one holding the live weights)
            for k in list(self._instances):
                if len(self._instances) < self.max_instances:
                    break
                victim = self._instances[k]
                if victim is self._primary or victim is self._current:
                    continue
                del self._instances[k]
                victim.graph = None
                del victim
            inst = _Instance(n, h, w, self.filters, self.dtype, self.use_tc)
            if init:
                inst.eng.init_params(self.seed)
            self._instances[key] = inst
        else:
            self._instances.move_to_end(key)
        return inst

    def _use(self, inst: _Instance):
        """Make `inst` hold the live weights (instances share nothing; parameters are copied flat)."""
        cur = self._current
        if inst is not cur:
            inst.eng.params.t.copy_(cur.eng.params.t)
            inst.eng.state.t.copy_(cur.eng.state.t)
            inst.eng.adam_m.copy_(cur.eng.adam_m)
            inst.eng.adam_v.copy_(cur.eng.adam_v)
            inst.eng.adam_state.copy_(cur.eng.adam_state)
            inst.eng._pack_dirty = True
            self._current = inst
        return inst

    @property
    def engine(self) -> Engine:
        return self._current.eng

    # ---- Keras weight protocol -----------------------------------------------------------------------
    def weight_names(self) -> List[str]:
        """Variable names in keras.Model.get_weights() order."""
        return self._primary.net.keras_weight_names()

    def creation_names(self) -> List[str]:
        return list(self._primary.net.creation_names)

    def get_weights(self) -> List[np.ndarray]:
        e = self._current.eng
        return [e.get_param(n) for n in self.weight_names()]

    def set_weights(self, weights: Sequence[np.ndarray]):
        names = self.weight_names()
        if len(weights) != len(names):
            raise ValueError(f"You called `set_weights(weights)` on a model with {len(names)} weights, "
                             f"but provided {len(weights)}")
        e = self._current.eng
        for n, w in zip(names, weights):
            e.set_param(n, _to_numpy(w))

    def get_named_weights(self) -> Dict[str, np.ndarray]:
        e = self._current.eng
        return {n: e.get_param(n) for n in self.creation_names()}

    def set_named_weights(self, named: Dict[str, np.ndarray]):
        e = self._current.eng
        for n in self.creation_names():
            e.set_param(n, _to_numpy(named[n]))

    def count_params(self) -> int:
        e = self._primary.eng
        return int(sum(np.prod(e.specs[n].logical_shape) for n in self.creation_names()))

    def to(self, device):
        if str(device).startswith("cpu"):
            raise L.SembError("this model runs on sm_100a only; there is no CPU path (use the reference for CPU inference)")
        return self

    # ---- inference -----------------------------------------------------------------------------------
    def __call__(self, x, training: bool = False):
        x = _to_numpy(x).astype(np.float32, copy=False)
        if x.ndim != 4 or x.shape[3] != 1:
            raise ValueError(f"expected NHWC input with one channel, got {x.shape}")
        inst = self._use(self._instance(x.shape[0], x.shape[1], x.shape[2]))
        e = inst.eng
        inst.x_pin.copy_(torch.from_numpy(np.ascontiguousarray(x)))
        inst.x_dev.copy_(inst.x_pin, non_blocking=True)
        if training:
            e.zero_step(zero_grads=False)
        inst.stage_in()
        e.forward(training=training)
        inst.stage_out()
        inst.out_pin.copy_(inst.out_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return inst.out_pin.clone()

    def predict(self, x, batch_size: int = 8):
        """Inference in chunks of exactly `batch_size` tiles (the last chunk is zero-padded: with moving statistics every
        sample is independent), so that one engine instance serves any number of tiles."""
        x = _to_numpy(x).astype(np.float32, copy=False)
        n = x.shape[0]
        batch_size = max(1, min(int(batch_size), n))
        outs = []
        for i in range(0, n, batch_size):
            chunk = x[i:i + batch_size]
            m = chunk.shape[0]
            if m < batch_size:
                chunk = np.concatenate([chunk, np.zeros((batch_size - m,) + chunk.shape[1:], dtype=np.float32)], 0)
            outs.append(self(chunk).numpy()[:m])
        return np.concatenate(outs, 0)

    def predict_tiled(self, img, tile_w: int, tile_h: int, min_overlap: int = 2, manage_overlap_mode: int = 2, batch_size: int = 16):
        """tile_image -> model -> stitch_image of the reference's inference loop (UNet_Segmentation.py:335-343) with the
        image uploaded ONCE: the overlapping tiles are gathered on the device straight into the engine's input, run in
        chunks of `batch_size` (the reference runs them one by one), and stitched on the device.  img: (H, W[, 1])."""
        from . import HelperFunctions as HF
        img = np.ascontiguousarray(_to_numpy(img), dtype=np.float32)
        if img.ndim == 3:
            img = img[:, :, 0]
        H, W = img.shape
        nx, xs = HF._grid(W, tile_w, min_overlap)
        ny, ys = HF._grid(H, tile_h, min_overlap)
        nt = nx * ny
        bs = max(1, min(int(batch_size), nt))
        inst = self._use(self._instance(bs, tile_h, tile_w))
        e = inst.eng
        dev = e.device
        img_d = torch.from_numpy(img).to(dev, non_blocking=True)
        xs_d = torch.tensor(xs, dtype=torch.int32, device=dev)
        ys_d = torch.tensor(ys, dtype=torch.int32, device=dev)
        pred = torch.zeros((nt, tile_h, tile_w), dtype=torch.float32, device=dev)
        out = torch.empty((H, W), dtype=torch.float32, device=dev)
        for k0 in range(0, nt, bs):
            cnt = min(bs, nt - k0)
            if cnt < bs:
                inst.x_dev.zero_()          # the padded tail of the last chunk (samples are independent in inference mode)
            L.check(e.lib.semb_tile_gather(img_d.data_ptr(), H, W, inst.x_dev.data_ptr(), tile_h, tile_w, xs_d.data_ptr(), nx,
                                           ys_d.data_ptr(), ny, k0, cnt, e.stream))
            inst.stage_in()
            e.forward(training=False)
            inst.stage_out()
            pred[k0:k0 + cnt].copy_(inst.out_dev[:cnt, :, :, 0])
        L.check(e.lib.semb_tile_stitch(pred.data_ptr(), tile_h, tile_w, out.data_ptr(), H, W, xs_d.data_ptr(), nx, ys_d.data_ptr(), ny,
                                       int(manage_overlap_mode), e.stream))
        return out.cpu().numpy()[:, :, None]

    # ---- training ------------------------------------------------------------------------------------
    def compile(self, weighting: float = 1.0, learning_rate: float = 1e-3, beta_1: float = 0.9, beta_2: float = 0.999,
                epsilon: float = 1e-7):
        self.weighting = float(weighting)
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon

    def set_distributed(self, process_group=None):
        """Data parallel over torch.distributed (NCCL): one flat all-reduce of the gradient buffer per step."""
        import torch.distributed as dist
        self.process_group = process_group
        self.world_size = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        from . import dp
        dp.broadcast_(self._current.eng.params.t, 0, process_group)
        dp.broadcast_(self._current.eng.state.t, 0, process_group)
        self._current.eng._pack_dirty = True

    def _step_device(self, inst: _Instance):
        """fwd + loss + bwd (+ all-reduce) + Adam on the device; inputs already in inst.x_dev / y_dev."""
        e = inst.eng
        e.lr.fill_(self.learning_rate)
        inst.stage_in()
        single = self.world_size == 1
        if self.use_cuda_graph:
            key = (self.weighting, self.beta_1, self.beta_2, self.epsilon, self.world_size)
            if inst.graph is None or inst.graph_key != key:
                # warm-up run on a side stream (allocates lazy gradient buffers), then capture
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                saved = [t.clone() for t in (e.params.t, e.state.t, e.adam_m, e.adam_v, e.adam_state)]
                with torch.cuda.stream(s):
                    inst.fwd_bwd(self.weighting)
                    if not single:
                        self._allreduce(e)      # also brings up the NCCL communicator before any capture
                    e.adam(self.beta_1, self.beta_2, self.epsilon, 1.0 / self.world_size)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                for t, sv in zip((e.params.t, e.state.t, e.adam_m, e.adam_v, e.adam_state), saved):
                    t.copy_(sv)
                e.repack()          # packed tensor-core weights follow the restored master weights
                g1, g2 = torch.cuda.CUDAGraph(), None
                one_graph = single or os.environ.get("SEMB_DP_TWO_GRAPHS") is None
                try:
                    # ONE replay per step: fwd + bwd, the NCCL all-reduce of the flat gradient buffer, Adam, weight re-pack
                    with torch.cuda.graph(g1):
                        inst.fwd_bwd(self.weighting)
                        if one_graph:
                            if not single:
                                self._allreduce(e)
                            e.adam(self.beta_1, self.beta_2, self.epsilon, 1.0 / self.world_size)
                except Exception:
                    if single:
                        raise
                    one_graph = False           # a collective that cannot be captured: all-reduce between two graphs
                    torch.cuda.synchronize()
                    g1 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g1):
                        inst.fwd_bwd(self.weighting)
                if not one_graph:
                    g2 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g2):
                        e.adam(self.beta_1, self.beta_2, self.epsilon, 1.0 / self.world_size)
                inst.graph, inst.graph2, inst.graph_key = g1, g2, key
            inst.graph.replay()
            if inst.graph2 is not None:
                self._allreduce(e)
                inst.graph2.replay()
        else:
            inst.fwd_bwd(self.weighting)
            if not single:
                self._allreduce(e)
            e.adam(self.beta_1, self.beta_2, self.epsilon, 1.0 / self.world_size)

    def _allreduce(self, e: Engine):
        from . import dp
        dp.allreduce_sum_(e.grads, self.process_group)

    def train_step(self, x, y) -> Dict[str, float]:
        """One TorchTrainer.train_step with HOST inputs: H2D, fwd, loss, bwd, Adam, D2H of the metrics."""
        if not isinstance(x, torch.Tensor):
            x = _to_numpy(x).astype(np.float32, copy=False)
        if not isinstance(y, torch.Tensor):
            y = _to_numpy(y).astype(np.float32, copy=False)
        inst = self._use(self._instance(x.shape[0], x.shape[1], x.shape[2]))
        io = inst.train_io()
        slot = io.upload(x, y)                              # copy stream: overlaps the previous step's graph
        main = torch.cuda.current_stream()
        main.wait_event(io.ev_h2d[slot])
        inst.x_dev.copy_(io.x_stage[slot], non_blocking=True)       # device copies; the graph reads x_dev / y_dev
        inst.y_dev.copy_(io.y_stage[slot], non_blocking=True)
        io.ev_free[slot].record(main)
        self._step_device(inst)
        r = io.step % io.RING
        io.step += 1
        io.sums_pin[r].copy_(inst.eng.zeroed.get(inst.loss_sums), non_blocking=True)
        io.ev_sums[r].record(main)
        return LazyLogs(io.sums_pin[r], io.ev_sums[r], float(inst.n * inst.h * inst.w))

    def _read_metrics(self, inst: _Instance) -> Dict[str, float]:
        e = inst.eng
        inst.sums_pin.copy_(e.zeroed.get(inst.loss_sums), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        cnt = float(inst.n * inst.h * inst.w)
        s = inst.sums_pin
        return {"loss": float(s[0]) / cnt, "mae": float(s[1]) / cnt, "acc": float(s[2]) / cnt}

    def test_step(self, x, y) -> Dict[str, float]:
        x = _to_numpy(x).astype(np.float32, copy=False)
        y = _to_numpy(y).astype(np.float32, copy=False)
        inst = self._use(self._instance(x.shape[0], x.shape[1], x.shape[2]))
        e = inst.eng
        inst.x_dev.copy_(torch.from_numpy(np.ascontiguousarray(x)))
        inst.y_dev.copy_(torch.from_numpy(np.ascontiguousarray(y)))
        e.zero_step(zero_grads=False)
        inst.stage_in()
        e.forward(training=False)
        inst.loss(self.weighting, with_grad=False)
        return self._read_metrics(inst)

    def fit(self, data, batch_size=None, epochs: int = 1, verbose: int = 1, callbacks=None, validation_data=None):
        """keras.Model.fit over a Sequence-like object (`__len__`, `__getitem__`, optional `on_epoch_end`)."""
        callbacks = list(callbacks or [])
        for cb in callbacks:
            cb.set_model(self)
        self.stop_training = False
        for cb in callbacks:
            cb.on_train_begin()
        for epoch in range(epochs):
            for cb in callbacks:
                cb.on_epoch_begin(epoch)
            t0 = time.time()
            agg: Dict[str, float] = {}
            nb = len(data)
            pending = collections.deque()
            for i in range(nb):
                xb, yb = data[i]
                pending.append(self.train_step(xb, yb))
                while len(pending) > (0 if i == nb - 1 else _TrainIO.RING - 2):      # read metrics a few steps late: no stall
                    for k, v in pending.popleft().items():
                        agg[k] = agg.get(k, 0.0) + v
            logs = {k: v / max(nb, 1) for k, v in agg.items()}
            if validation_data is not None and len(validation_data) > 0:
                vagg: Dict[str, float] = {}
                for i in range(len(validation_data)):
                    xb, yb = validation_data[i]
                    for k, v in self.test_step(xb, yb).items():
                        vagg[k] = vagg.get(k, 0.0) + v
                logs.update({"val_" + k: v / len(validation_data) for k, v in vagg.items()})
            logs["learning_rate"] = self.learning_rate
            if verbose and self.rank == 0:
                msg = " - ".join(f"{k}: {v:.4f}" for k, v in logs.items())
                print(f"Epoch {epoch + 1}/{epochs} - {time.time() - t0:.1f}s - {msg}", flush=True)
            self.history.append(dict(logs, epoch=epoch))
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs)
            if hasattr(data, "on_epoch_end"):
                data.on_epoch_end()
            if self.stop_training:
                break
        for cb in callbacks:
            cb.on_train_end()
        return self.history

    # ---- persistence ---------------------------------------------------------------------------------
    def save(self, path: str):
        """`model.save(path)` (UNet_Segmentation.py:262,287): a Keras-3 style `.keras` zip (metadata.json, config.json,
        model.weights.npz store with Keras layer paths; see keras_io) -- or, for a path ending in .npz, the flat archive of
        round 1."""
        named = self.get_named_weights()
        cfg = {"class": "MultiResUNet", "input_shape": list(self.input_shape), "filters": self.filters,
               "dtype": self.dtype, "weighting": self.weighting, "format": "semb200-keras-1"}
        if path.endswith(".npz"):
            with open(path, "wb") as fh:
                np.savez(fh, __config__=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8), **named)
            return
        from . import keras_io
        keras_io.save_keras(path, cfg, named, self.creation_names())

    @staticmethod
    def load(path: str, dtype: Optional[str] = None, batch_size: int = 1, input_shape=None, filters: int = 16) -> "UNetModel":
        """`.keras` (this package's writer), legacy `.npz`, or one of the reference's frozen `.pb` UNets
        (ImageJ Plugin/SEM_Particle_Segmentation_Models; input_shape / filters must then be given, default 512x352x1 / 16)."""
        from . import keras_io
        if path.endswith(".pb"):
            m = UNetModel(tuple(input_shape or (512, 352, 1)), filters, 1, dtype or "f32", batch_size)
            m.set_named_weights(keras_io.read_pb_weights(path, m.creation_names()))
            return m
        if keras_io.is_keras_archive(path):
            cfg, named = keras_io.load_keras(path)
        else:
            with np.load(path) as z:
                cfg = json.loads(bytes(z["__config__"]).decode())
                named = {k: z[k] for k in z.files if k != "__config__"}
        m = UNetModel(tuple(input_shape or cfg["input_shape"]), cfg["filters"], 1, dtype or cfg["dtype"], batch_size)
        m.set_named_weights(named)
        m.weighting = cfg.get("weighting", 1.0)
        return m


def load_model(path: str, custom_objects=None, **kw) -> UNetModel:
    """keras.models.load_model stand-in (custom_objects accepted and ignored, UNet_Segmentation.py:303)."""
    return UNetModel.load(path, **kw)


# ---- callbacks (the three the reference uses: UNet_Segmentation.py:262-272) --------------------------------
class Callback:
    model: UNetModel = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self): pass
    def on_train_end(self): pass
    def on_epoch_begin(self, epoch): pass
    def on_epoch_end(self, epoch, logs): pass


class ModelCheckpoint(Callback):
    def __init__(self, filepath, monitor="loss", verbose=0, save_best_only=False, mode="min"):
        self.filepath, self.monitor, self.verbose, self.save_best_only = filepath, monitor, verbose, save_best_only
        self.best = math.inf if mode == "min" else -math.inf
        self.mode = mode

    def on_epoch_end(self, epoch, logs):
        if self.model.rank != 0:
            return
        path = self.filepath.format(epoch=epoch + 1)
        cur = logs.get(self.monitor)
        if self.save_best_only:
            better = cur is not None and (cur < self.best if self.mode == "min" else cur > self.best)
            if not better:
                return
            self.best = cur
        if self.verbose:
            print(f"Epoch {epoch + 1}: saving model to {path}")
        self.model.save(path)


class CSVLogger(Callback):
    def __init__(self, filename, separator=",", append=False):
        self.filename, self.sep, self.append = filename, separator, append
        self.keys = None

    def on_epoch_end(self, epoch, logs):
        if self.model.rank != 0:
            return
        new = not (self.append and os.path.exists(self.filename)) and self.keys is None
        if self.keys is None:
            self.keys = sorted(logs.keys())
        with open(self.filename, "w" if new else "a") as fh:
            if new:
                fh.write(self.sep.join(["epoch"] + self.keys) + "\n")
            fh.write(self.sep.join([str(epoch)] + [repr(float(logs[k])) for k in self.keys]) + "\n")


class LearningRateScheduler(Callback):
    def __init__(self, schedule):
        self.schedule = schedule

    def on_epoch_begin(self, epoch):
        self.model.learning_rate = float(self.schedule(epoch, self.model.learning_rate))
