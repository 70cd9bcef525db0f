"""ctypes binding of libss2d_b200.so — the only way the Python side reaches the kernels.

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is
raised.  The structs mirror include/ss2d_b200.h field by field.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libss2d_b200.so")

F32, F16, BF16 = 0, 1, 2
FAMILY_AUTO, FAMILY_STATELANES, FAMILY_WARPSCAN = 0, 1, 2
REF_CHUNK = 2048
SL_BLOCK = 16  # sequences up to this length need no checkpoints at all

_vp, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int32


class ScanFwdParams(C.Structure):
    _fields_ = (
        [(n, _i64) for n in ("batch", "dim", "seqlen", "dstate", "ngroups")]
        + [(n, _i32) for n in ("in_dtype", "out_dtype", "delta_softplus", "family")]
        + [(n, _vp) for n in ("u", "delta", "A", "B", "C", "D", "delta_bias", "z")]
        + [(n, _i64) for n in ("u_bstride", "u_dstride", "delta_bstride", "delta_dstride", "B_bstride", "B_gstride",
                                "B_nstride", "C_bstride", "C_gstride", "C_nstride", "z_bstride", "z_dstride")]
        + [("out", _vp), ("out_bstride", _i64), ("out_dstride", _i64), ("out_z", _vp), ("x", _vp), ("ckpt", _vp)]
    )


class ScanBwdParams(C.Structure):
    _fields_ = (
        [("f", ScanFwdParams), ("dout", _vp), ("dout_bstride", _i64), ("dout_dstride", _i64), ("ckpt_scratch", _vp)]
        + [(n, _vp) for n in ("du", "ddelta", "dz", "dA", "dB", "dC", "dD", "ddelta_bias")]
    )


class CrossFwdParams(C.Structure):
    _fields_ = (
        [(n, _i64) for n in ("batch", "D", "H", "W", "dstate")]
        + [("in_dtype", _i32), ("delta_softplus", _i32), ("family", _i32), ("deterministic", _i32)]
        + [(n, _vp) for n in ("x", "delta", "B", "C", "A", "Dskip", "delta_bias", "y", "ckpt")]
        + [("bc_bstride", _i64), ("bc_gstride", _i64), ("work", _vp)]
    )


class CrossBwdParams(C.Structure):
    _fields_ = [("f", CrossFwdParams)] + [(n, _vp) for n in ("dy", "ckpt_scratch", "dx", "ddelta", "dA", "dB", "dC",
                                                             "dDskip", "ddelta_bias")]


EXPORTS = ("ss2d_abi_version", "ss2d_build_info", "ss2d_error_string", "ss2d_scan_ckpt_floats", "ss2d_cross_work_floats",
           "ss2d_scan_family", "ss2d_set_default_family", "ss2d_cross_family", "ss2d_optim_partials", "ss2d_optim_clip_adam", "ss2d_dt_proj_fwd", "ss2d_dt_proj_bwd",
           "ss2d_plane_transpose", "ss2d_selective_scan_fwd",
           "ss2d_selective_scan_bwd", "ss2d_cross_scan", "ss2d_cross_merge", "ss2d_cross_scan_fwd",
           "ss2d_cross_scan_bwd", "ss2d_dwconv_silu_fwd", "ss2d_dwconv_silu_bwd", "ss2d_merge_norm_gate_fwd",
           "ss2d_merge_norm_gate_bwd", "ss2d_cross_permute")

_lib = None


def lib():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m focalnet_b200.build` "
                "(focalnet_b200 has no CPU or PyTorch fallback path)")
        L = C.CDLL(LIB_PATH)
        L.ss2d_abi_version.restype = C.c_int
        L.ss2d_build_info.restype = C.c_char_p
        L.ss2d_error_string.restype = C.c_char_p
        L.ss2d_error_string.argtypes = [C.c_int]
        L.ss2d_scan_ckpt_floats.restype = _i64
        L.ss2d_scan_ckpt_floats.argtypes = [_i64, _i64, _i64, _i64]
        L.ss2d_cross_work_floats.restype = _i64
        L.ss2d_cross_work_floats.argtypes = [_i64, _i64, _i64, _i64, _i64, _i32, _i32, _i32]
        L.ss2d_scan_family.restype = C.c_int
        L.ss2d_scan_family.argtypes = [_vp]
        L.ss2d_set_default_family.restype = C.c_int
        L.ss2d_set_default_family.argtypes = [C.c_int]
        L.ss2d_cross_family.restype = C.c_int
        L.ss2d_cross_family.argtypes = [_i64, _i64, _i64, _i64, _i64, _i32, _i32]
        L.ss2d_optim_partials.restype = _i64
        L.ss2d_optim_partials.argtypes = []
        _f = C.c_float
        L.ss2d_optim_clip_adam.restype = C.c_int
        L.ss2d_optim_clip_adam.argtypes = [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _f, _f, _f, _f, _i64, _f, _f, _vp]
        sigs = {n: [_vp, _vp] for n in ("ss2d_selective_scan_fwd", "ss2d_selective_scan_bwd", "ss2d_cross_scan_fwd",
                                        "ss2d_cross_scan_bwd")}
        sigs.update({n: [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp] for n in ("ss2d_cross_scan", "ss2d_cross_merge")})
        sigs["ss2d_dwconv_silu_fwd"] = [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp]
        sigs["ss2d_dwconv_silu_bwd"] = [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _vp]
        sigs["ss2d_plane_transpose"] = [_vp, _vp, _i64, _i64, _i64, _i32, _vp]
        sigs["ss2d_cross_permute"] = [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _vp]
        sigs["ss2d_dt_proj_fwd"] = [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp]
        sigs["ss2d_dt_proj_bwd"] = [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp]
        _f32 = C.c_float
        sigs["ss2d_merge_norm_gate_fwd"] = [_vp, _vp, _vp, _f32, _vp, _i64, _vp, _i64, _i64, _i64, _vp]
        sigs["ss2d_merge_norm_gate_bwd"] = [_vp, _vp, _vp, _f32, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _vp]
        for name, argtypes in sigs.items():
            fn = getattr(L, name)  # AttributeError here == a symbol of include/ss2d_b200.h is not exported
            fn.restype = C.c_int
            fn.argtypes = argtypes
        if L.ss2d_abi_version() != 3:
            raise RuntimeError("libss2d_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().ss2d_error_string(rc).decode()} (code {rc})")
