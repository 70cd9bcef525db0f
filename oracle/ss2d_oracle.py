"""TEST INFRASTRUCTURE ONLY — CPU oracle for the SS2D hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``focalnet_b200``) never does: it has no CPU path.

Two things live here:

* ctypes bindings of ``oracle/ss2d_oracle.c`` (double-precision plain-C restatement; the checker
  used by the parity tests).  Built by ``oracle/build_oracle.sh`` / ``__graft_entry__.build()``.
* ``selective_scan_ref_port`` — a PyTorch port with the same *algorithmic structure* as the
  reference's CPU path ``selective_scan_ref`` (/root/reference/kernels/selective_scan/
  test_selective_scan.py:168-234: materialise exp(delta*A) and delta*B*u as (B,D,L,N) tensors, then a
  Python loop over L).  It is what ``bench.py`` times as the CPU baseline (kind "port").

Pinned against the reference itself by tests/test_oracle_golden.py (vectors from oracle/make_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libss2d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "ss2d_oracle.c")
    ):
        subprocess.check_call(["bash", os.path.join(_HERE, "build_oracle.sh")])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ss2d_oracle_scan_fwd.restype = ctypes.c_int
        _lib.ss2d_oracle_scan_bwd.restype = ctypes.c_int
    return _lib


def _f32(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        t = t.detach().float().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def _p(a):
    return ctypes.c_void_p(0) if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _i(v):
    return ctypes.c_int64(int(v))


def _lift_bc(B):
    return B[:, None] if B.ndim == 3 else B


def scan_fwd(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False):
    """-> dict(out (B,Dm,L), x (B,Dm,nchunks,2N), last_state (B,Dm,N)) as float32 numpy."""
    u, delta, A, B, C, D, z, delta_bias = map(_f32, (u, delta, A, B, C, D, z, delta_bias))
    B, C = _lift_bc(B), _lift_bc(C)
    Bn, Dm, L = u.shape
    N, G = A.shape[1], B.shape[1]
    nch = (L + 2047) // 2048
    out = np.empty((Bn, Dm, L), np.float32)
    x = np.empty((Bn, Dm, nch, 2 * N), np.float32)
    last = np.empty((Bn, Dm, N), np.float32)
    rc = lib().ss2d_oracle_scan_fwd(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(z), _p(delta_bias),
                                    ctypes.c_int(int(delta_softplus)), _i(Bn), _i(Dm), _i(L), _i(N), _i(G),
                                    _p(out), _p(x), _p(last))
    if rc != 0:
        raise RuntimeError(f"ss2d_oracle_scan_fwd failed: {rc}")
    return dict(out=out, x=x, last_state=last)


def scan_bwd(u, delta, A, B, C, D, z, delta_bias, dout, delta_softplus=False):
    """-> dict(du, ddelta, dA, dB, dC, dD, ddelta_bias, dz) as float32 numpy (None where input None)."""
    u, delta, A, B, C, D, z, delta_bias, dout = map(_f32, (u, delta, A, B, C, D, z, delta_bias, dout))
    squeeze = B.ndim == 3
    B, C = _lift_bc(B), _lift_bc(C)
    Bn, Dm, L = u.shape
    N, G = A.shape[1], B.shape[1]
    du, dd = np.empty_like(u), np.empty_like(u)
    dA = np.empty_like(A)
    dB, dC = np.empty_like(B), np.empty_like(C)
    dD = np.empty(Dm, np.float32) if D is not None else None
    db = np.empty(Dm, np.float32) if delta_bias is not None else None
    dz = np.empty_like(u) if z is not None else None
    rc = lib().ss2d_oracle_scan_bwd(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(z), _p(delta_bias), _p(dout),
                                    ctypes.c_int(int(delta_softplus)), _i(Bn), _i(Dm), _i(L), _i(N), _i(G),
                                    _p(du), _p(dd), _p(dA), _p(dB), _p(dC), _p(dD), _p(db), _p(dz))
    if rc != 0:
        raise RuntimeError(f"ss2d_oracle_scan_bwd failed: {rc}")
    if squeeze:
        dB, dC = dB[:, 0], dC[:, 0]
    return dict(du=du, ddelta=dd, dA=dA, dB=dB, dC=dC, dD=dD, ddelta_bias=db, dz=dz)


def cross_scan(x):
    x = _f32(x)
    B, C, H, W = x.shape
    xs = np.empty((B, 4, C, H * W), np.float32)
    lib().ss2d_oracle_cross_scan(_p(x), _p(xs), _i(B), _i(C), _i(H), _i(W))
    return xs


def cross_merge(ys):
    ys = _f32(ys)
    B, K, C, H, W = ys.shape
    assert K == 4
    y = np.empty((B, C, H * W), np.float32)
    lib().ss2d_oracle_cross_merge(_p(ys), _p(y), _i(B), _i(C), _i(H), _i(W))
    return y


def dwconv_silu_fwd(xin, weight, bias, C=None):
    """xin:(B,H,W,Cs) channels-last (first C channels are convolved); weight:(C,1,3,3)|(C,3,3)."""
    xin, weight, bias = _f32(xin), _f32(weight), _f32(bias)
    B, H, W, Cs = xin.shape
    C = C or weight.shape[0]
    out = np.empty((B, C, H, W), np.float32)
    lib().ss2d_oracle_dwconv_silu_fwd(_p(xin), _i(Cs), _p(weight.reshape(C, 9)), _p(bias), _p(out),
                                      _i(B), _i(C), _i(H), _i(W))
    return out


def dwconv_silu_bwd(xin, weight, bias, dout, C=None):
    xin, weight, bias, dout = _f32(xin), _f32(weight), _f32(bias), _f32(dout)
    B, H, W, Cs = xin.shape
    C = C or weight.shape[0]
    dx = np.empty((B, H, W, C), np.float32)
    dw = np.empty((C, 9), np.float32)
    db = np.empty(C, np.float32) if bias is not None else None
    lib().ss2d_oracle_dwconv_silu_bwd(_p(xin), _i(Cs), _p(weight.reshape(C, 9)), _p(bias), _p(dout), _p(dx),
                                      _p(dw), _p(db), _i(B), _i(C), _i(H), _i(W))
    return dict(dx=dx, dweight=dw.reshape(weight.shape), dbias=db)


# ----------------------------------------------------------------------------------------------
# PyTorch port of the reference CPU path (what bench.py times as cpu_baseline, kind="port").
# ----------------------------------------------------------------------------------------------
def selective_scan_ref_port(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                            return_last_state=False, compute_dtype=torch.float32):
    """Same signature/semantics as the reference's ``selective_scan_ref``
    (test_selective_scan.py:168-234), real-valued A only.  Differentiable (autograd)."""
    in_dtype = u.dtype
    u = u.to(compute_dtype)
    dt = delta.to(compute_dtype)
    if delta_bias is not None:
        dt = dt + delta_bias.to(compute_dtype).unsqueeze(-1)
    if delta_softplus:
        dt = F.softplus(dt)
    A = A.to(compute_dtype)
    Bn, Dm, L = u.shape
    N = A.shape[1]
    Bv, Cv = B.to(compute_dtype), C.to(compute_dtype)
    if Bv.dim() == 3:
        Bv = Bv.unsqueeze(1)
    if Cv.dim() == 3:
        Cv = Cv.unsqueeze(1)
    rep = Dm // Bv.shape[1]
    Bv = Bv.repeat_interleave(rep, dim=1)          # (B, Dm, N, L)
    Cv = Cv.repeat_interleave(rep, dim=1)
    decay = torch.exp(dt.unsqueeze(-1) * A.view(1, Dm, 1, N))                    # (B,Dm,L,N)
    drive = (dt * u).unsqueeze(-1) * Bv.transpose(2, 3)                          # (B,Dm,L,N)
    h = u.new_zeros((Bn, Dm, N))
    ys = []
    for t in range(L):
        h = decay[:, :, t] * h + drive[:, :, t]
        ys.append((h * Cv[:, :, :, t]).sum(-1))
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + u * D.to(compute_dtype).view(1, Dm, 1)
    if z is not None:
        y = y * F.silu(z.to(compute_dtype))
    y = y.to(in_dtype)
    return (y, h) if return_last_state else y


def cross_scan_port(x):
    """(B,C,H,W)->(B,4,C,L); torch twin of CrossScan.forward (vmamba_layers.py:31-38)."""
    rowmajor = x.flatten(2)
    colmajor = x.transpose(2, 3).flatten(2)
    fwd = torch.stack([rowmajor, colmajor], dim=1)
    return torch.cat([fwd, fwd.flip(-1)], dim=1)


def cross_merge_port(ys):
    """(B,4,C,H,W)->(B,C,L); torch twin of CrossMerge.forward (vmamba_layers.py:52-58)."""
    B, K, C, H, W = ys.shape
    ys = ys.reshape(B, K, C, H * W)
    pair = ys[:, :2] + ys[:, 2:].flip(-1)
    return pair[:, 0] + pair[:, 1].reshape(B, C, W, H).transpose(2, 3).reshape(B, C, H * W)
