"""ctypes binding of libsemb200.so (the C ABI declared in include/semb200.h).

The product path has NO fallback: if the shared library cannot be loaded (or built from
csrc/ with nvcc) every entry point raises, and on a machine with a GPU `require_device()`
raises unless the device is sm_100.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

OK = 0
F32, BF16 = 0, 1
PAD_ZERO, PAD_REFLECT = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4
AFF_NONE, AFF_PLAIN, AFF_BATCH = 0, 1, 2

_ERR_NAMES = {-1: "SEMB_ESHAPE", -2: "SEMB_EALIGN", -3: "SEMB_EARCH", -4: "SEMB_EWORKSPACE", -5: "SEMB_ECUDA"}


class Tensor(C.Structure):
    """semb_tensor"""
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("pitch", C.c_int32), ("coff", C.c_int32)]


class ConvGeom(C.Structure):
    """semb_conv_geom"""
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "OH", "OW", "Cin", "Cout", "R", "S", "stride", "pad_t", "pad_l",
                                         "pad_mode", "dtype")]


class AffineDesc(C.Structure):
    """semb_affine_desc"""
    _fields_ = [(n, C.c_int32) for n in ("N", "HW", "C", "dtype", "act", "actb", "mode_a", "mode_b", "aff_nstride", "nseg_b")] + \
               [("seg_c0", C.c_int32 * 3), ("seg_b", Tensor * 3), ("seg_db", Tensor * 3)]


class NormFin(C.Structure):
    """semb_norm_fin"""
    _fields_ = [("stats", C.c_void_p), ("stats_nstride", C.c_int32), ("cstride", C.c_int32), ("count", C.c_float), ("eps", C.c_float),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p),
                ("invstd", C.c_void_p), ("moving_mean", C.c_void_p), ("moving_var", C.c_void_p), ("momentum", C.c_float)]


_P = C.c_void_p
_I = C.c_int32
_L = C.c_int64
_F = C.c_float
_TP = C.POINTER(Tensor)
_GP = C.POINTER(ConvGeom)
_AP = C.POINTER(AffineDesc)

# name -> (restype, argtypes); must list every symbol declared in include/semb200.h
SIGNATURES = {
    "semb_version": (C.c_int, []),
    "semb_last_error": (C.c_char_p, []),
    "semb_launch_count": (C.c_int64, []),
    "semb_device_ok": (C.c_int, []),
    "semb_conv2d_fwd": (C.c_int, [_GP, _TP, _P, _P, _TP, _P, _I, _I, _I, _P]),
    "semb_conv2d_dgrad": (C.c_int, [_GP, _TP, _P, _P, _TP, _P, _I, _I, _I, _P]),
    "semb_conv2d_wgrad": (C.c_int, [_GP, _TP, _TP, _P, _P, _P]),
    "semb_pack_weights_tc": (C.c_int64, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "semb_pack_batch_job_size": (C.c_int64, []),
    "semb_pack_batch_prepare": (C.c_int, [_I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "semb_pack_weights_tc_batch": (C.c_int, [_P, _I, _I, _P]),
    "semb_conv2d_fwd_tc": (C.c_int, [_GP, _TP, _P, _P, _TP, _P, _I, _I, _I, _P]),
    "semb_conv2d_wgrad_tc": (C.c_int, [_GP, _TP, _TP, _P, _P]),
    "semb_conv2d_wgrad_tc_workspace": (C.c_int64, [_GP]),
    "semb_conv2d_wgrad_tc_ws": (C.c_int, [_GP, _TP, _TP, _P, _P, _L, _P]),
    "semb_norm_finalize": (C.c_int, [_P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P]),
    "semb_norm_from_moving": (C.c_int, [_I, _F, _P, _P, _P, _P, _P, _P, _P]),
    "semb_affine_act_fwd": (C.c_int, [_AP, _TP, _P, _P, _TP, _P, _P, _TP, _P, _I, _I, _P]),
    "semb_affine_act_bwd_reduce": (C.c_int, [_AP, _TP, _TP, _TP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "semb_norm_bwd_finalize": (C.c_int, [_P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P]),
    "semb_affine_act_bwd_apply": (C.c_int, [_AP, _TP, _TP, _TP, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                            _TP, _I, _TP, _I, _P]),
    "semb_affine_act_fwd_fin": (C.c_int, [_AP, _TP, C.POINTER(NormFin), _TP, C.POINTER(NormFin), _TP, _P, _I, _I, _P]),
    "semb_affine_act_bwd_apply_sums": (C.c_int, [_AP, _TP, _TP, _TP, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _F, _P, _P,
                                                 _P, _I, _I, _TP, _I, _TP, _I, _P]),
    "semb_affine_act_bwd_fused": (C.c_int, [_AP, _TP, _TP, _TP, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _F, _P, _P,
                                            _P, _I, _I, _P, _TP, _I, _TP, _I, _P]),
    "semb_channel_sum": (C.c_int, [_TP, _I, _I, _P, _I, _P]),
    "semb_maxpool2x2_fwd": (C.c_int, [_TP, _TP, _I, _I, _I, _I, _P]),
    "semb_maxpool2x2_bwd": (C.c_int, [_TP, _TP, _TP, _I, _I, _I, _I, _I, _P]),
    "semb_pad_crop": (C.c_int, [_TP, _TP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "semb_pixel_shuffle2": (C.c_int, [_TP, _TP, _I, _I, _I, _P, _I, _I, _P]),
    "semb_pixel_shuffle2x": (C.c_int, [_TP, _TP, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    "semb_split_bf16": (C.c_int, [_TP, _TP, _L, _P]),
    "semb_split_weights": (C.c_int, [_P, _I, _I, _I, _I, _P, _I, _I, _P]),
    "semb_conv2d_fwd_tc_f32": (C.c_int, [_GP, _TP, _P, _P, _TP, _P, _I, _I, _I, _P]),
    "semb_conv2d_fwd_tc_d2s": (C.c_int, [_GP, _TP, _P, _P, _TP, _I, _I, _P]),
    "semb_mask_mul": (C.c_int, [_TP, _TP, _TP, _TP, C.c_int64, C.c_float, _I, _I, _P]),
    "semb_gp_direction": (C.c_int, [_TP, _TP, _I, C.c_int64, C.c_float, _P, _I, _P]),
    "semb_tile_gather": (C.c_int, [_P, _I, _I, _P, _I, _I, _P, _I, _P, _I, _I, _I, _P]),
    "semb_tile_stitch": (C.c_int, [_P, _I, _I, _P, _I, _I, _P, _I, _P, _I, _I, _P]),
    "semb_upsample2x": (C.c_int, [_TP, _TP, _I, _I, _I, _I, _I, _I, _P]),
    "semb_s2d_weights": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "semb_merge_weights": (C.c_int, [_P, _P, _I, _I, _I, _P, _I, _P]),
    "semb_fold_stats4": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "semb_tap_patch": (C.c_int, [_TP, _TP, _I, _I, _I, _I, _P, _I, _I, _P]),
    "semb_tapfold_weights": (C.c_int, [_P, _I, _I, _I, _P, _I, _I, _P]),
    "semb_loss_wbce": (C.c_int, [_TP, _P, _TP, _L, _F, _P, _I, _P]),
    "semb_loss_wbce_logits": (C.c_int, [_TP, _P, _P, _P, _TP, _L, _F, _P, _I, _P]),
    "semb_loss_l1_l2": (C.c_int, [_TP, _TP, _F, _I, _L, _I, _F, _TP, _I, _P, _I, _P]),
    "semb_adam_step": (C.c_int, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _P, _P]),
    "semb_fill_f32": (C.c_int, [_P, _L, _F, _P]),
    "semb_cast_in": (C.c_int, [_P, _I, _TP, _L, _I, _P]),
    "semb_cast_out": (C.c_int, [_TP, _P, _I, _L, _I, _P]),
}

_lib = None


class SembError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """dlopen libsemb200.so, building it first when only the sources are present."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SEMB_LIB_PATH") or _build.LIB_PATH      # override: A/B builds of the kernels (tuning only)
    if not os.path.exists(path):
        path = _build.build()  # raises when nvcc is unavailable: no fallback path exists
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library is stale: rebuild
        fn.restype = res
        fn.argtypes = args
    if lib.semb_version() != 100:
        raise SembError(f"libsemb200.so version {lib.semb_version()} does not match the binding (100)")
    _lib = lib
    return lib


def last_error() -> str:
    return load().semb_last_error().decode()


def check(rc: int):
    """status code -> exception (shape/alignment problems are ValueError like Keras', the rest RuntimeError)."""
    if rc == OK:
        return
    msg = f"{_ERR_NAMES.get(rc, rc)}: {last_error()}"
    if rc in (-1, -2):
        raise ValueError(msg)
    raise SembError(msg)


def require_device():
    import torch
    if not torch.cuda.is_available():
        raise SembError("libsemb200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if not load().semb_device_ok():
        raise SembError(last_error())


def launch_count() -> int:
    return int(load().semb_launch_count())
