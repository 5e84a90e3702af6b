"""ctypes binding of libddpm_b200.so (the C-ABI declared in include/ddpm_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libddpm_b200.so"

MAX_TAPS = 9
ABI_VERSION = 3

_ll = C.c_longlong
_vp = C.c_void_p
_i = C.c_int
_f = C.c_float
_ull = C.c_ulonglong


class ConvArgs(C.Structure):
    _fields_ = [
        ("x0", _vp), ("c0", _i), ("ld0", _ll),
        ("x1", _vp), ("c1", _i), ("ld1", _ll),
        ("n", _i), ("h", _i), ("w", _i),
        ("src_n", _i),
        ("ntaps", _i),
        ("tap_dn", _i * MAX_TAPS), ("tap_dh", _i * MAX_TAPS), ("tap_dw", _i * MAX_TAPS), ("tap_wk", _i * MAX_TAPS),
        ("wgt", _vp), ("cout", _i), ("ldw", _ll), ("k_total", _ll),
        ("out", _vp), ("out_f32", _vp), ("ldo", _ll),
        ("bias", _vp),
        ("temb", _vp), ("ld_temb", _i),
        ("res", _vp), ("ldr", _ll),
        ("gn_x0", _vp), ("gn_ld0", _ll), ("gn_c0", _i),
        ("gn_x1", _vp), ("gn_ld1", _ll),
        ("gn_coef", _vp),
        ("gn_silu", _i),
        ("gn_sums", _vp),
        ("out_csum", _vp),
        ("splitk_ws", _vp), ("splitk_ws_elems", _ll),
        ("split_io", _i),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("dy", _vp), ("ldy", _ll), ("cout", _i),
        ("x0", _vp), ("c0", _i), ("ld0", _ll),
        ("x1", _vp), ("c1", _i), ("ld1", _ll),
        ("n", _i), ("h", _i), ("w", _i), ("src_n", _i),
        ("ntaps", _i),
        ("tap_dn", _i * MAX_TAPS), ("tap_dh", _i * MAX_TAPS), ("tap_dw", _i * MAX_TAPS), ("tap_wk", _i * MAX_TAPS),
        ("dw", _vp), ("ldw", _ll),
        ("accumulate", _i),
        ("splits", _i),
        ("dbias", _vp),
    ]


class PrepDesc(C.Structure):
    _fields_ = [
        ("w", _vp), ("wf", _vp), ("ldwf", _ll), ("wd", _vp), ("ldwd", _ll),
        ("cout", _i), ("taps", _i), ("cin", _i),
        ("tile_begin", _i), ("tiles_x", _i), ("tiles_y", _i),
    ]


# name -> argtypes (every function returns int except ddpm_last_error)
SIGNATURES = {
    "ddpm_abi_version": [],
    "ddpm_add_noise": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _vp],
    "ddpm_mse_fwd_bwd": [_vp, _vp, _vp, _vp, _ll, _vp],
    "ddpm_scale_by_device_scalar": [_vp, _vp, _ll, _vp],
    "ddpm_bgemm": [_vp, _ll, _ll, _ll, _i, _vp, _ll, _ll, _ll, _i, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _i, _f, _vp],
    "ddpm_softmax_rows": [_vp, _ll, _vp, _ll, _ll, _i, _vp],
    "ddpm_softmax_rows_bwd": [_vp, _ll, _vp, _ll, _vp, _ll, _i, _f, _vp],
    "ddpm_resize_h_u8": [_vp, _vp, _ll, _i, _i, _i, _vp, _vp, _i, _vp],
    "ddpm_resize_v_normalize": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp],
    "ddpm_gn_bwd_dparams": [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp],
    "ddpm_sumsq_f32": [_vp, _ll, _vp, _vp],
    "ddpm_adamw_flat": [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _f, _vp, _f, _f, _f, _f, _f, _vp],
    "ddpm_scheduler_step": [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _f, _vp],
    "ddpm_ddim_step": [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _f, _i, _vp],
    "ddpm_scheduler_step_philox": [_vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _f, _ull, _ull, _vp],
    "ddpm_to_uint8_nhwc": [_vp, _vp, _i, _i, _i, _i, _vp],
    "ddpm_unipc_x0": [_vp, _vp, _vp, _ll, _f, _f, _vp],
    "ddpm_unipc_update": [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _f, _vp],
    "ddpm_conv_gemm": [C.POINTER(ConvArgs), _vp],
    "ddpm_conv_gemm_workspace_elems": [C.POINTER(ConvArgs)],
    "ddpm_conv_halo_strips": [_i],
    "ddpm_conv_wgrad": [C.POINTER(WgradArgs), _vp],
    "ddpm_prep_weight": [_vp, _vp, _ll, _vp, _ll, _i, _i, _i, _vp],
    "ddpm_prep_weights_batched": [_vp, _i, _i, _i, _vp],
    "ddpm_im2col3": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "ddpm_nhwc_to_nchw_f32": [_vp, _ll, _vp, _i, _i, _i, _i, _vp],
    "ddpm_gn_stats": [_vp, _i, _ll, _vp, _i, _ll, _i, _i, _i, _vp, _vp],
    "ddpm_gn_apply": [_vp, _i, _ll, _vp, _i, _ll, _i, _i, _i, _vp, _f, _vp, _vp, _i, _vp, _ll, _vp, _vp],
    "ddpm_gn_stats_from_csum": [_vp, _i, _vp, _i, _i, _i, _vp, _vp],
    "ddpm_gn_stats_split": [_vp, _i, _ll, _vp, _i, _ll, _i, _i, _i, _vp, _vp],
    "ddpm_gn_apply_split": [_vp, _i, _ll, _vp, _i, _ll, _i, _i, _i, _vp, _f, _vp, _vp, _i, _vp, _ll, _vp],
    "ddpm_attn_fwd_split": [_vp, _ll, _vp, _ll, _i, _i, _i, _i, _f, _vp],
    "ddpm_im2col3_split": [_vp, _vp, _i, _i, _i, _i, _vp],
    "ddpm_gn_fwd": [_vp, _i, _ll, _vp, _i, _ll, _i, _i, _i, _f, _vp, _vp, _i, _vp, _vp, _ll, _vp, _vp, _vp],
    "ddpm_gn_bwd": [_vp, _i, _ll, _vp, _i, _ll, _i, _i, _i, _vp, _f, _vp, _vp, _i, _vp, _ll, _vp, _ll, _vp, _ll,
                    _vp, _ll, _vp, _ll, _vp, _vp, _vp, _vp],
    "ddpm_gn_bwd_apply": [_vp, _i, _ll, _vp, _i, _ll, _i, _i, _i, _vp, _f, _vp, _vp, _ll, _vp, _vp, _ll, _vp, _ll,
                          _vp, _ll, _vp, _ll, _vp, _vp, _vp, _ll, _vp, _vp],
    "ddpm_attn_fwd": [_vp, _ll, _vp, _ll, _vp, _i, _i, _i, _i, _f, _vp],
    "ddpm_attn_bwd": [_vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _i, _i, _i, _i, _f, _vp],
    "ddpm_attn_wide_supported": [_i, _i, _i],
    "ddpm_attn_wide_fwd": [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _f, _vp],
    "ddpm_timestep_embedding": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "ddpm_linear_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "ddpm_linear_f32_wgrad": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "ddpm_linear_f32_dgrad": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "ddpm_reduce_hw": [_vp, _ll, _i, _i, _i, _vp, _ll, _vp, _vp],
    "ddpm_dropout": [_vp, _vp, _vp, _ll, _f, _ull, _ull, _vp, _vp],
    "ddpm_space_to_depth": [_vp, _ll, _vp, _i, _i, _i, _i, _i, _vp],
    "ddpm_zero_insert2x": [_vp, _ll, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "ddpm_upsample2x": [_vp, _ll, _vp, _i, _i, _i, _i, _vp],
    "ddpm_sumpool2x": [_vp, _ll, _vp, _ll, _vp, _i, _i, _i, _i, _vp],
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("DDPM_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"{path} not found: build it with `python -m polyp_image_generator_b200.build` "
            "(there is no CPU or eager fallback for the DDPM hot path)")
    lib = C.CDLL(str(path))
    lib.ddpm_last_error.restype = C.c_char_p
    lib.ddpm_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = C.c_longlong if name.endswith("_workspace_elems") else C.c_int
        fn.argtypes = argtypes
    if lib.ddpm_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libddpm_b200.so ABI {lib.ddpm_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def last_error() -> str:
    return load().ddpm_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")
