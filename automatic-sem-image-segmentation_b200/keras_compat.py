"""The sliver of the Keras functional API that the reference's static layer functions are written against
(UNet_Segmentation.py:401-503, :565-589): `Input`, symbolic tensors with `.shape`, `Model(inputs, outputs)` with
`model(x, training)`, `get_weights / set_weights`, and the `ReflectionPadding2D` layer class.

    inputs = keras_compat.Input(shape=(128, 128, 1), batch_size=4, dtype="bf16")
    x = UNet.multi_res_block(16, inputs)
    x = UNet.res_path(16, 2, x)
    y = UNet.conv2d_bn(x, 1, 1, 1, activation="sigmoid")
    model = keras_compat.Model(inputs, y)

Every call records engine ops (hand-written CUDA behind the C ABI); nothing here computes.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib as L
from .engine import Engine, PadCropOp
from .nets import T, UNetBuilder


def Input(shape, batch_size: int = 1, dtype: str = "bf16", name=None) -> T:
    """keras.layers.Input(shape=(H, W, C)): opens a new graph (engine) specialised to `batch_size`."""
    h, w = int(shape[0]), int(shape[1])
    c = int(shape[2]) if len(shape) > 2 else 1
    eng = Engine(batch_size, dtype)
    b = UNetBuilder(eng, h, w, in_channels=c, build=False)
    return b.input


def builder_of(t: T) -> UNetBuilder:
    b = getattr(t.view.buf.eng, "_builder", None)
    if b is None:
        raise TypeError("expected a symbolic tensor created from keras_compat.Input(...)")
    return b


def shape_of(t: T):
    return (None, t.h, t.w, t.layout.logical)


T.shape = property(shape_of)          # the reference's functions read `tensor.shape[...]` (:335, :518-519)


class ReflectionPadding2D:
    """UNet_Segmentation.py:565-589 / CycleGAN.py:482-506: `padding=(width_total, height_total)`, split total//2 before and
    total//2 + total%2 after, np.pad(mode='reflect') semantics."""

    def __init__(self, padding=(2, 2), **kwargs):
        self.padding = tuple(padding)

    def __call__(self, x: T, mask=None) -> T:
        pw, ph = self.padding
        if pw == 0 and ph == 0:
            return x
        b = builder_of(x)
        e = b.e
        out = e.new_buf(x.h + ph, x.w + pw, x.layout.phys, b.anon("reflection_padding2d"), requires_grad=x.view.requires_grad)
        e.add_op(PadCropOp(e, x.view, out.view(), (x.h, x.w), (x.h + ph, x.w + pw), ph // 2, pw // 2, "reflect"))
        return T(out.view(), x.h + ph, x.w + pw, x.layout, b.kg.layer("reflection_padding2d", [x.klayer]))

    def get_config(self):
        return {"padding": self.padding}


class Model:
    """keras.models.Model(inputs, outputs) over a recorded graph: forward in either mode, Keras-ordered weights."""

    def __init__(self, inputs: T, outputs: T, name=None):
        self.b = builder_of(outputs)
        if builder_of(inputs) is not self.b:
            raise ValueError("inputs and outputs belong to different graphs")
        self.b.set_output(outputs)
        self.eng = self.b.e
        self.eng.finalize()
        self.eng.init_params(0)
        self.inputs, self.outputs = inputs, outputs
        n, h, w, c = self.eng.N, inputs.h, inputs.w, inputs.layout.logical
        self._x = torch.zeros((n, h, w, c), dtype=torch.float32, device=self.eng.device)
        self._y = torch.zeros((n, outputs.h, outputs.w, outputs.layout.logical), dtype=torch.float32, device=self.eng.device)

    def weight_names(self) -> List[str]:
        return self.b.keras_weight_names()

    def get_weights(self) -> List[np.ndarray]:
        return [self.eng.get_param(n) for n in self.weight_names()]

    def set_weights(self, weights: Sequence[np.ndarray]):
        names = self.weight_names()
        if len(weights) != len(names):
            raise ValueError(f"You called `set_weights(weights)` on a model with {len(names)} weights, but provided {len(weights)}")
        for n, w in zip(names, weights):
            self.eng.set_param(n, np.asarray(w))

    def count_params(self) -> int:
        return int(sum(np.prod(self.eng.specs[n].logical_shape) for n in self.b.creation_names))

    def __call__(self, x, training: bool = False) -> torch.Tensor:
        e = self.eng
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
        if tuple(x.shape) != tuple(self._x.shape):
            raise ValueError(f"this graph was recorded for inputs of shape {tuple(self._x.shape)}, got {tuple(x.shape)}")
        self._x.copy_(torch.from_numpy(x))
        cin, cout = self._x.shape[3], self._y.shape[3]
        L.check(e.lib.semb_cast_in(self._x.data_ptr(), cin, C.byref(self.b.in_buf.view().t), x.shape[0] * x.shape[1] * x.shape[2], e.dtype, e.stream))
        e.zero_step(zero_grads=False)
        e.forward(training=training)
        ov = self.outputs.view
        L.check(e.lib.semb_cast_out(C.byref(ov.t), self._y.data_ptr(), cout, self._y.shape[0] * self._y.shape[1] * self._y.shape[2], e.dtype,
                                    e.stream))
        return self._y.cpu()


class Activation:
    """keras.layers.Activation(name) as the reference passes it to its block functions (CycleGAN.py:378-389)."""

    def __init__(self, name: str, **kwargs):
        codes = {"relu": L.ACT_RELU, "tanh": L.ACT_TANH, "sigmoid": L.ACT_SIGMOID, "linear": L.ACT_NONE}
        if name not in codes:
            raise NotImplementedError(f"activation {name!r} has no fused kernel")
        self.name, self.act = name, codes[name]


class LeakyReLU:
    """keras.layers.LeakyReLU(negative_slope): the kernels implement the slope the reference uses (0.2)."""

    def __init__(self, negative_slope: float = 0.2, **kwargs):
        if abs(negative_slope - 0.2) > 1e-12:
            raise NotImplementedError("only LeakyReLU(0.2) (CycleGAN.py:436-447) is built")
        self.act = L.ACT_LEAKY


def act_code(activation) -> int:
    if activation is None:
        return L.ACT_NONE
    if isinstance(activation, str):
        return Activation(activation).act
    return activation.act
