"""On-disk model containers.

`.keras` (UNet_Segmentation.py:262,287,303; CycleGAN.py:203,221,228): Keras 3 writes a zip archive with `metadata.json`,
`config.json` and a weights store.  h5py is not available here, so the store written is Keras' OTHER supported variant,
`model.weights.npz` (keras.src.saving.saving_lib.NpzIOStore: one entry per layer path `layers/<layer name>/vars`, each a
dict {"0": array, "1": array, ...} in `layer.weights` order), with Keras' automatic layer names (`conv2d`, `conv2d_1`,
..., `batch_normalization_7`, `conv2d_transpose_3`).  `config.json` carries this package's own model description (class,
input shape, filters, options) -- NOT a serialised Keras functional graph, so Keras can read the weights store of such a
file but not rebuild the model from it; that direction is UNPINNED (Keras cannot be installed in this environment).
Reading accepts what this module writes, plus the legacy `.npz` files of round 1.

`.pb` (ImageJ Plugin/SEM_Particle_Segmentation_Models/*.pb, frozen TF-1.12 GraphDefs): `read_pb_weights` extracts the
MultiRes-UNet variables by name with a raw protobuf wire-format reader, so the reference's shipped weights load without
TensorFlow.
"""
from __future__ import annotations

import io
import json
import re
import struct
import time
import zipfile
from typing import Dict, List, Tuple

import numpy as np

KERAS_VERSION = "3.5.0"


def keras_layer_name(tf1_name: str) -> str:
    """`conv2d_1` (TF-1 / creation numbering used for the variable names here) -> `conv2d`, `conv2d_2` -> `conv2d_1`, ...
    Names without a numeric suffix (the CycleGAN layers) are kept."""
    m = re.fullmatch(r"(.*)_(\d+)", tf1_name)
    if not m:
        return tf1_name
    k = int(m.group(2))
    return m.group(1) if k == 1 else f"{m.group(1)}_{k - 1}"


def group_by_layer(names: List[str]) -> List[Tuple[str, List[str]]]:
    """creation-order variable names `layer/var` -> [(layer, [variables in layer.weights order])]"""
    out: List[Tuple[str, List[str]]] = []
    for n in names:
        layer = n.rsplit("/", 1)[0]
        if out and out[-1][0] == layer:
            out[-1][1].append(n)
        else:
            out.append((layer, [n]))
    return out


def save_keras(path: str, config: dict, named: Dict[str, np.ndarray], order: List[str], rename=keras_layer_name):
    store = {}
    for layer, vs in group_by_layer(order):
        store[f"layers/{rename(layer)}/vars"] = np.array({str(i): np.asarray(named[v]) for i, v in enumerate(vs)}, dtype=object)
    buf = io.BytesIO()
    np.savez(buf, **store)
    meta = {"keras_version": KERAS_VERSION, "date_saved": time.strftime("%Y-%m-%d@%H:%M:%S"), "writer": "sem_b200"}
    cfg = dict(config, variable_order=order)
    with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
        z.writestr("metadata.json", json.dumps(meta))
        z.writestr("config.json", json.dumps(cfg))
        z.writestr("model.weights.npz", buf.getvalue())


def load_keras(path: str, rename=keras_layer_name) -> Tuple[dict, Dict[str, np.ndarray]]:
    with zipfile.ZipFile(path) as z:
        cfg = json.loads(z.read("config.json"))
        names = set(z.namelist())
        if "model.weights.npz" not in names:
            raise ValueError(f"{path}: only the model.weights.npz store can be read here (model.weights.h5 needs h5py)")
        with np.load(io.BytesIO(z.read("model.weights.npz")), allow_pickle=True) as w:
            store = {k: w[k].item() for k in w.files}
    named = {}
    for layer, vs in group_by_layer(cfg["variable_order"]):
        d = store[f"layers/{rename(layer)}/vars"]
        for i, v in enumerate(vs):
            named[v] = np.asarray(d[str(i)])
    return cfg, named


def is_keras_archive(path: str) -> bool:
    return zipfile.is_zipfile(path) and "config.json" in zipfile.ZipFile(path).namelist()


# ---- frozen TF-1.12 GraphDef reader (raw protobuf wire format) ----------------------------------------------------------
def _varint(b: bytes, i: int):
    r, s = 0, 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        if c < 0x80:
            return r, i
        s += 7


def _fields(b: bytes):
    i, n = 0, len(b)
    while i < n:
        key, i = _varint(b, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 1:
            v, i = b[i:i + 8], i + 8
        elif wt == 2:
            ln, i = _varint(b, i)
            v, i = b[i:i + ln], i + ln
        elif wt == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


def _tensor(b: bytes) -> np.ndarray:
    shape, content, fvals = [], None, []
    for fn, wt, v in _fields(b):
        if fn == 2:                                   # TensorShapeProto
            for f2, _, d in _fields(v):
                if f2 == 2:                           # Dim
                    size = 0
                    for f3, _, s in _fields(d):
                        if f3 == 1:
                            size = s
                    shape.append(size)
        elif fn == 4:
            content = v
        elif fn == 5:                                 # float_val, packed or repeated
            fvals.extend(struct.unpack(f"<{len(v) // 4}f", v) if wt == 2 else struct.unpack("<f", v))
    n = int(np.prod(shape)) if shape else 1
    if content is not None:
        return np.frombuffer(content, dtype="<f4").reshape(shape).copy()
    if len(fvals) == 1:
        return np.full(shape, fvals[0], dtype=np.float32)
    return np.asarray(fvals, dtype=np.float32).reshape(shape) if len(fvals) == n else np.zeros(shape, dtype=np.float32)


def read_pb_consts(path: str) -> Dict[str, np.ndarray]:
    """{node name: float tensor} of every Const node of a frozen GraphDef."""
    with open(path, "rb") as fh:
        data = fh.read()
    out = {}
    for fn, _, node in _fields(data):
        if fn != 1:
            continue
        name, op, tensor = None, None, None
        for f2, _, v in _fields(node):
            if f2 == 1:
                name = v.decode()
            elif f2 == 2:
                op = v.decode()
            elif f2 == 5:                             # attr map entry {key, value}
                key, val = None, None
                for f3, _, e in _fields(v):
                    if f3 == 1:
                        key = e.decode()
                    elif f3 == 2:
                        val = e
                if key == "value" and val is not None:
                    for f4, _, t in _fields(val):
                        if f4 == 8:
                            tensor = t
        if op == "Const" and tensor is not None:
            try:
                out[name] = _tensor(tensor)
            except Exception:
                pass
    return out


def read_pb_weights(path: str, variable_names: List[str]) -> Dict[str, np.ndarray]:
    """The MultiRes-UNet variables of a frozen reference graph: node `conv2d_7_1/kernel` <-> variable `conv2d_7/kernel`."""
    consts = read_pb_consts(path)
    named = {}
    for v in variable_names:
        layer, var = v.rsplit("/", 1)
        key = f"{layer}_1/{var}"
        if key not in consts:
            raise KeyError(f"{path}: no Const node {key}")
        named[v] = consts[key]
    return named
