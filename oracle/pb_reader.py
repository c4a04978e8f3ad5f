"""Minimal TensorFlow GraphDef (protobuf wire format) reader.  TEST INFRASTRUCTURE.

Extracts the Const tensors of the reference's frozen TF-1.12 UNet graphs
(ImageJ Plugin/SEM_Particle_Segmentation_Models/TiO2_UNet_Masks_*.pb) without
TensorFlow.  Field numbers:
  GraphDef.node=1; NodeDef{name=1, op=2, input=3, attr=5(map entry key=1,value=2)};
  AttrValue{tensor=8}; TensorProto{dtype=1, tensor_shape=2, tensor_content=4, float_val=5};
  TensorShapeProto.dim=2; Dim.size=1.
"""
from __future__ import annotations

import struct
import numpy as np


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fn, wt, v


def _parse_shape(buf):
    dims = []
    for fn, wt, v in _fields(buf):
        if fn == 2:
            size = 0
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _parse_tensor(buf):
    dtype, shape, content, fvals = None, (), None, []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            dtype = v
        elif fn == 2:
            shape = _parse_shape(v)
        elif fn == 4:
            content = bytes(v)
        elif fn == 5:
            if wt == 2:
                fvals.extend(struct.unpack(f"<{len(v) // 4}f", bytes(v)))
            else:
                fvals.append(struct.unpack("<f", bytes(v))[0])
    if dtype != 1:  # DT_FLOAT
        return None
    n = int(np.prod(shape)) if shape else 1
    if content:
        arr = np.frombuffer(content, dtype="<f4").copy()
    elif len(fvals) == 1:
        arr = np.full((n,), fvals[0], dtype=np.float32)
    else:
        arr = np.asarray(fvals, dtype=np.float32)
    return arr.reshape(shape)


def read_graph_consts(path):
    """Return ({node_name: np.ndarray} for float Const nodes, [(name, op)] node list)."""
    with open(path, "rb") as f:
        data = memoryview(f.read())
    consts, nodes = {}, []
    for fn, wt, v in _fields(data):
        if fn != 1:
            continue
        name, op, tensor = None, None, None
        for f2, w2, v2 in _fields(v):
            if f2 == 1:
                name = bytes(v2).decode()
            elif f2 == 2:
                op = bytes(v2).decode()
            elif f2 == 5:
                key, val = None, None
                for f3, w3, v3 in _fields(v2):
                    if f3 == 1:
                        key = bytes(v3).decode()
                    elif f3 == 2:
                        val = v3
                if key == "value" and val is not None:
                    for f4, w4, v4 in _fields(val):
                        if f4 == 8:
                            tensor = _parse_tensor(v4)
        nodes.append((name, op))
        if op == "Const" and tensor is not None:
            consts[name] = tensor
    return consts, nodes


def unet_params_from_pb(path):
    """Map the .pb Const nodes onto oracle.unet.UNetSpec names (creation order ==
    the '_N_1' numbering of the frozen graph)."""
    consts, _ = read_graph_consts(path)
    out = {}
    for name, arr in consts.items():
        parts = name.split("/")
        if len(parts) != 2:
            continue
        layer, var = parts
        if not layer.endswith("_1"):
            continue
        base = layer[:-2]
        if base.startswith(("conv2d_transpose_", "conv2d_", "batch_normalization_")):
            out[f"{base}/{var}"] = arr
    return out
