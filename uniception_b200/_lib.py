"""ctypes binding of `libuc_b200.so` (C ABI: include/uc_b200.h).

The shared library is the product; there is NO fallback.  If it is missing this module raises at
import (build it with `python __graft_entry__.py` or `make -C uniception_b200/csrc`), and every
compute entry point returns an error on a machine without an sm_100 GPU, which `check()` turns
into `RuntimeError` (the reference's native op raises RuntimeError through TORCH_CHECK,
curope.cpp:54-59).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuc_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
        "Build it with `python __graft_entry__.py` or `make -C uniception_b200/csrc`."
    )

lib = C.CDLL(LIB_PATH)

UC_DTYPE_BF16, UC_DTYPE_F32, UC_DTYPE_F16 = 0, 1, 2
EPI_BIAS, EPI_ROPE, EPI_GELU, EPI_GELU_BWD, EPI_RESIDUAL, EPI_ATOMIC, EPI_RELU, EPI_RELU_BWD, EPI_RESIDUAL_F32 = 1, 2, 4, 8, 16, 32, 64, 128, 256

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmParams(C.Structure):
    _fields_ = [
        ("a", vp), ("b", vp), ("c", vp),
        ("m", i32), ("n", i32), ("k", i32),
        ("a_layout", i32), ("b_layout", i32),
        ("lda", i64), ("ldb", i64), ("ldc", i64),
        ("c_dtype", i32), ("epilogue", i32), ("split_k", i32), ("rope_cols", i32),
        ("bias", vp), ("residual", vp), ("aux_out", vp), ("aux_in", vp), ("positions", vp), ("rope_table", vp),
        ("c_colsum", vp),
    ]


class Rope2dParams(C.Structure):
    _fields_ = [
        ("tokens", vp), ("positions", vp),
        ("B", i32), ("N", i32), ("H", i32), ("D", i32),
        ("stride_b", i64), ("stride_n", i64), ("stride_h", i64),
        ("dtype", i32), ("base", f32), ("fwd", f32),
    ]


class LayerNormFwdParams(C.Structure):
    _fields_ = [
        ("x", vp), ("y", vp), ("gamma", vp), ("beta", vp), ("mean", vp), ("rstd", vp),
        ("rows", i32), ("C", i32), ("x_dtype", i32), ("y_dtype", i32), ("eps", f32),
    ]


class LayerNormBwdParams(C.Structure):
    _fields_ = [
        ("dy", vp), ("x", vp), ("dres", vp), ("dx", vp), ("gamma", vp), ("mean", vp), ("rstd", vp),
        ("dgamma", vp), ("dbeta", vp),
        ("rows", i32), ("C", i32), ("dy_dtype", i32), ("x_dtype", i32),
        ("dx_colsum", vp),
    ]


class AttnFwdParams(C.Structure):
    _fields_ = [
        ("q", vp), ("k", vp), ("v", vp), ("o", vp), ("lse", vp),
        ("B", i32), ("H", i32), ("Nq", i32), ("Nk", i32),
        ("ldq", i64), ("ldk", i64), ("ldv", i64), ("ldo", i64),
        ("scale", f32),
    ]


class AttnBwdParams(C.Structure):
    _fields_ = [
        ("q", vp), ("k", vp), ("v", vp), ("o", vp), ("d_o", vp), ("lse", vp), ("delta", vp), ("dq_acc", vp),
        ("dq", vp), ("dk", vp), ("dv", vp),
        ("B", i32), ("H", i32), ("Nq", i32), ("Nk", i32),
        ("ldq", i64), ("ldk", i64), ("ldv", i64), ("ldo", i64), ("lddq", i64), ("lddk", i64), ("lddv", i64),
        ("scale", f32),
        ("q_positions", vp), ("k_positions", vp), ("rope_table", vp),
    ]


class HeadPostFwdParams(C.Structure):
    _fields_ = [
        ("y", vp), ("pts", vp), ("conf", vp),
        ("B", i32), ("h", i32), ("w", i32), ("patch", i32),
        ("conf_min", f32), ("conf_max", f32), ("ldy", i64),
    ]


class HeadPostBwdParams(C.Structure):
    _fields_ = [
        ("y", vp), ("dpts", vp), ("dconf", vp), ("dy", vp), ("dy_dtype", i32),
        ("B", i32), ("h", i32), ("w", i32), ("patch", i32),
        ("conf_min", f32), ("conf_max", f32), ("ldy", i64),
    ]


class HeadNormParams(C.Structure):
    _fields_ = [
        ("x", vp), ("y", vp), ("ldx", i64), ("ldy", i64), ("gamma", vp), ("beta", vp), ("dgamma", vp), ("dbeta", vp),
        ("positions", vp), ("rope_table", vp), ("rows", i32), ("heads", i32), ("eps", f32),
    ]


class PatchEmbedParams(C.Structure):
    _fields_ = [("img", vp), ("w", vp), ("bias", vp), ("out", vp), ("B", i32), ("H", i32), ("W", i32), ("patch", i32), ("n", i32)]


class Conv3x3Params(C.Structure):
    _fields_ = [
        ("mode", i32), ("B", i32), ("H", i32), ("W", i32), ("cin", i32), ("cout", i32),
        ("x", vp), ("w", vp), ("y", vp), ("dy", vp), ("dx", vp), ("dw", vp), ("bias", vp), ("residual", vp), ("relu_out", vp),
        ("relu", i32),
    ]


# every symbol include/uc_b200.h declares (checked by tests/test_cabi.py)
EXPORTS = {
    "uc_version": (C.c_int, []),
    "uc_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "uc_launch_count": (C.c_uint64, []),
    "uc_gemm": (C.c_int, [C.POINTER(GemmParams), vp]),
    "uc_set_gemm_dynamic": (C.c_int, [C.c_int]),
    "uc_rope2d": (C.c_int, [C.POINTER(Rope2dParams), vp]),
    "uc_rope2d_table": (C.c_int, [vp, i32, i32, f32, f32, vp]),
    "uc_layernorm_fwd": (C.c_int, [C.POINTER(LayerNormFwdParams), vp]),
    "uc_layernorm_bwd": (C.c_int, [C.POINTER(LayerNormBwdParams), vp]),
    "uc_attn_fwd": (C.c_int, [C.POINTER(AttnFwdParams), vp]),
    "uc_attn_bwd": (C.c_int, [C.POINTER(AttnBwdParams), vp]),
    "uc_patchify": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "uc_colsum": (C.c_int, [vp, i32, i64, i32, i32, vp, vp]),
    "uc_cast_bf16": (C.c_int, [vp, vp, i64, vp]),
    "uc_nlc_to_nchw": (C.c_int, [vp, i32, vp, i32, i32, i32, vp]),
    "uc_nchw_to_nlc": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
    "uc_head_post_fwd": (C.c_int, [C.POINTER(HeadPostFwdParams), vp]),
    "uc_head_post_bwd": (C.c_int, [C.POINTER(HeadPostBwdParams), vp]),
    "uc_softmax_rows_fwd": (C.c_int, [vp, vp, i32, i32, i32, f32, vp]),
    "uc_softmax_rows_bwd": (C.c_int, [vp, vp, vp, i32, i32, i32, f32, vp]),
    "uc_patch_embed": (C.c_int, [C.POINTER(PatchEmbedParams), vp]),
    "uc_conv3x3": (C.c_int, [C.POINTER(Conv3x3Params), vp]),
    "uc_im2col3x3": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "uc_col2im3x3": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "uc_depth_space": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "uc_bilinear_fwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "uc_bilinear_fwd_f32in": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "uc_bilinear_bwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "uc_elementwise": (C.c_int, [i32, vp, vp, vp, vp, i64, vp]),
    "uc_headnorm_fwd": (C.c_int, [C.POINTER(HeadNormParams), vp]),
    "uc_headnorm_bwd": (C.c_int, [C.POINTER(HeadNormParams), vp]),
    "uc_layerscale_fwd": (C.c_int, [vp, vp, vp, vp, i32, i32, vp]),
    "uc_layerscale_bwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, vp]),
}

for _name, (_res, _args) in EXPORTS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib.uc_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a UC_ERR_* return code to the exception type the reference raises for the same failure:
    shape / argument problems are AssertionError-like in the Python reference but RuntimeError in
    its native op (TORCH_CHECK) -- the native convention is kept for everything below the C ABI."""
    if rc != 0:
        raise RuntimeError(f"libuc_b200 error {rc}: {last_error()}")


def launch_count() -> int:
    return int(lib.uc_launch_count())
