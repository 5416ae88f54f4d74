"""ctypes binding of libm2d_b200.so (C ABI declared in include/m2d.h).

There is no fallback: if the shared library is missing or the device is not
sm_100-class, importing callers get a RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libm2d_b200.so")

ACT = {"id": 0, "none": 0, None: 0, "relu": 1, "leaky": 2, "tanh": 3}

f32p = C.c_void_p      # device pointers travel as integers
i64 = C.c_longlong


class RowConvArgs(C.Structure):
    _fields_ = [
        ("x", f32p), ("x_bs", i64), ("x_ld", C.c_int), ("x_rows", C.c_int),
        ("nb", C.c_int),
        ("win_T", C.c_int), ("win_stride", C.c_int), ("win_pad", C.c_int), ("win_seq_len", C.c_int),
        ("w", f32p), ("w_ld", C.c_int),
        ("w_tiled", f32p),
        ("N", C.c_int), ("T", C.c_int), ("Cc", C.c_int),
        ("sr", C.c_int), ("roff0", C.c_int), ("droff", C.c_int),
        ("y", f32p), ("y_bs", i64), ("y_ld", C.c_int), ("y_rows", C.c_int),
        ("y2", f32p),
        ("bias", f32p),
        ("act", C.c_int),
        ("mask", f32p), ("m_bs", i64), ("m_ld", C.c_int), ("mask_mode", C.c_int),
        ("add", f32p), ("a_bs", i64), ("a_ld", C.c_int), ("add_before_mask", C.c_int),
        ("ws", f32p), ("ws_floats", i64),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("dy", f32p), ("dy_bs", i64), ("dy_ld", C.c_int), ("dy_rows", C.c_int),
        ("nb", C.c_int),
        ("x", f32p), ("x_bs", i64), ("x_ld", C.c_int), ("x_rows", C.c_int),
        ("win_T", C.c_int), ("win_stride", C.c_int), ("win_pad", C.c_int), ("win_seq_len", C.c_int),
        ("Cout", C.c_int), ("T", C.c_int), ("Cc", C.c_int),
        ("sr", C.c_int), ("roff0", C.c_int), ("droff", C.c_int),
        ("dw", f32p),
        ("packed", C.c_int),
        ("scale", C.c_float), ("beta", C.c_float),
        ("ws", f32p), ("ws_floats", i64),
    ]


_I, _F, _P, _L = C.c_int, C.c_float, C.c_void_p, i64

# name -> argtypes (every entry returns int unless listed in _RESTYPE)
SIGNATURES = {
    "m2d_version": [],
    "m2d_check_device": [_I],
    "m2d_set_gemm_mode": [_I],
    "m2d_get_gemm_mode": [],
    "m2d_halo_launch_count": [],
    "m2d_halo_persist_launch_count": [],
    "m2d_rowconv": [C.POINTER(RowConvArgs), _P],
    "m2d_wgrad": [C.POINTER(WgradArgs), _P],
    "m2d_wgrad_min_ws": [_I, _I, _I],
    "m2d_pack_conv_fwd": [_P, _P, _I, _I, _I, _P],
    "m2d_pack_conv_bwd": [_P, _P, _I, _I, _I, _I, _P],
    "m2d_pack_batch": [_P, _I, _P],
    "m2d_conv_dgrad_c1": [_P, _I, _I, _I, _P, _I, _I, _I, _P, _I, _P],
    "m2d_set_gru_impl": [_I],
    "m2d_set_gru_forward_batch_group": [_I],
    "m2d_gru_forward": [_P, _P, _P, _P, _I, _P, _I, _I, _I, _P],
    "m2d_gru_backward": [_P, _I, _P, _I, _P, _P, _P, _P, _I, _I, _I, _P],
    "m2d_colstats": [_P, _I, _L, _I, _P, _P],
    "m2d_bn_apply": [_P, _I, _P, _I, _L, _I, _P, _P, _P, _P, _P, _F, _F, _P, _I, _P],
    "m2d_colstats_groups": [_P, _I, _L, _I, _I, _P, _P],
    "m2d_bn_apply_groups": [_P, _I, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P, _F, _F, _P, _I, _P],
    "m2d_bn_train": [_P, _I, _P, _I, _L, _I, _P, _P, _P, _P, _P, _F, _F, _P, _I, _P],
    "m2d_bn_eval": [_P, _I, _P, _I, _L, _I, _P, _P, _P, _P, _F, _I, _P],
    "m2d_bn_bwd_reduce": [_P, _I, _P, _I, _P, _I, _L, _I, _P, _I, _P, _P],
    "m2d_bn_bwd_apply": [_P, _I, _P, _I, _P, _I, _P, _I, _L, _I, _P, _P, _I, _P, _P, _P, _P],
    "m2d_colsum": [_P, _I, _L, _I, _P, _F, _F, _P, _P],
    "m2d_axpby": [_P, _P, _P, _L, _F, _F, _P],
    "m2d_fill": [_P, _L, _F, _P],
    "m2d_scale_rows": [_P, _P, _P, _I, _L, _P],
    "m2d_interp": [_P, _P, _P, _P, _I, _L, _P],
    "m2d_rows_sumsq": [_P, _I, _L, _P, _P],
    "m2d_sum": [_P, _L, _P, _P],
    "m2d_gp_finalize": [_P, _P, _I, _P, _P, _P, _F, _P],
    "m2d_interp_stack3": [_P, _P, _P, _P, _I, _L, _P],
    "m2d_colsum_batch": [_P, _I, _I, _P, _P],
    "m2d_gp_finalize_lp": [_P, _I, _P, _P, _P],
    "m2d_pose_losses": [_P, _P, _P, _I, _I, _I, _F, _F, _I, _P, _P],
    "m2d_jerkiness": [_P, _I, _I, _I, _P, _P],
    "m2d_act_bwd": [_P, _P, _L, _I, _P],
    "m2d_crop_batch": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "m2d_mul3": [_P, _I, _P, _I, _P, _I, _P, _I, _L, _I, _F, _P],
    "m2d_maxpool2": [_P, _I, _P, _I, _I, _I, _I, _P],
    "m2d_maxpool2_bwd": [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P],
    "m2d_upsample2": [_P, _I, _P, _I, _I, _I, _I, _P],
    "m2d_upsample2_bwd": [_P, _I, _P, _I, _I, _I, _I, _I, _P],
    "m2d_copy2d": [_P, _I, _P, _I, _L, _I, _I, _P],
    "m2d_copy2d_batch": [_P, _I, _P],
    "m2d_fusion_mlp": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "m2d_embed_rows": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "m2d_embed_grad": [_P, _I, _P, _P, _I, _I, _I, _I, _F, _F, _P],
    "m2d_transpose_bcl": [_P, _P, _I, _I, _I, _P],
    "m2d_wgan_scalars": [_P, _P, _I, _L, _L, _F, _F, _I, _P, _P, _P, _P],
    "m2d_slice_audio": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "m2d_adam": [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _P],
    "m2d_adam_pack": [_P, _I, _I, _P, _F, _F, _F, _F, _F, _P],
    "m2d_nvl_allreduce": [_P, _P, _P, _I, _I, _L, _L, _I, _I, _P, _P],
    "m2d_nvl_allreduce2": [_P, _P, _L, _L, _P, _P, _L, _L, _P, _I, _I, _I, _I, _P, _P],
    "m2d_timestamp": [_P, _P],
}
_RESTYPE = {"m2d_wgrad_min_ws": i64, "m2d_halo_launch_count": i64, "m2d_halo_persist_launch_count": i64}
_NOCHECK = {"m2d_wgrad_min_ws", "m2d_version", "m2d_get_gemm_mode", "m2d_halo_launch_count",
            "m2d_halo_persist_launch_count"}     # return a value, not a status

_lib = None


def load():
    """Load the shared library once; raise loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(music2dance_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.m2d_last_error.restype = C.c_char_p
    lib.m2d_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is missing
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().m2d_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libm2d_b200 {what} failed (code {rc}): {msg}")


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if name not in _NOCHECK and rc != 0:
        check(rc, name)
    return rc
