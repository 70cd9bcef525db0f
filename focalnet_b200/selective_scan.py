"""Host-side mirror of the reference's selective-scan operator interface (seam S1), on top of the C ABI.

Mirrors, with the same names / argument meaning / error behaviour:
  * ``fwd`` / ``bwd`` of the pybind modules ``selective_scan_cuda_oflex`` and ``selective_scan_cuda_core``
    (kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_oflex.cpp:157-363,
    cus/selective_scan.cpp) — see ``focalnet_b200/dropin/`` for the importable modules;
  * ``selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)`` — the API
    of record (kernels/selective_scan/test_selective_scan.py:18-165).

torch is used only for allocation, strides and the current stream; the arithmetic is in libss2d_b200.so.
There is no CPU / eager fallback: non-CUDA tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


def _chk(cond: bool, msg: str):
    if not cond:
        raise RuntimeError(msg)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_scan_inputs(u, delta, A, B, C, D, delta_bias, z=None):
    """The TORCH_CHECKs of selective_scan_oflex.cpp:166-216 (same conditions, RuntimeError on violation)."""
    _chk(u.dtype in _DT, "selective_scan: u must be float32, float16 or bfloat16")
    _chk(A.dtype == torch.float32, "selective_scan: A must be float32")
    for name, t in (("delta", delta), ("B", B), ("C", C)):
        _chk(t.dtype == u.dtype, f"selective_scan: {name} must have the dtype of u")
    for name, t in (("u", u), ("delta", delta), ("A", A), ("B", B), ("C", C)):
        _chk(t.is_cuda, f"selective_scan: {name} must be a CUDA tensor (there is no CPU path)")
    _chk(u.dim() == 3 and B.dim() == 4 and C.dim() == 4 and A.dim() == 2, "selective_scan: bad ranks")
    batch, dim, L = u.shape
    N, G = A.shape[1], B.shape[1]
    _chk(u.stride(-1) == 1 or L == 1, "selective_scan: u.stride(-1) must be 1")
    _chk(delta.stride(-1) == 1 or L == 1, "selective_scan: delta.stride(-1) must be 1")
    _chk(dim % G == 0, "dims should be dividable by n_groups")
    _chk(N <= 256, "selective_scan only supports state dimension <= 256")
    _chk(tuple(delta.shape) == (batch, dim, L), "selective_scan: delta must have shape (batch, dim, seqlen)")
    _chk(tuple(A.shape) == (dim, N), "selective_scan: A must have shape (dim, dstate)")
    _chk(A.is_contiguous(), "selective_scan: A must be contiguous")
    _chk(tuple(B.shape) == (batch, G, N, L), "selective_scan: B must have shape (batch, n_groups, dstate, seqlen)")
    _chk(tuple(C.shape) == (batch, G, N, L), "selective_scan: C must have shape (batch, n_groups, dstate, seqlen)")
    _chk(B.stride(-1) == 1 or L == 1, "selective_scan: B.stride(-1) must be 1")
    _chk(C.stride(-1) == 1 or L == 1, "selective_scan: C.stride(-1) must be 1")
    for name, t in (("D", D), ("delta_bias", delta_bias)):
        if t is not None:
            _chk(t.dtype == torch.float32 and t.is_cuda, f"selective_scan: {name} must be a float32 CUDA tensor")
            _chk(tuple(t.shape) == (dim,) and t.is_contiguous(), f"selective_scan: {name} must have shape (dim,)")
    if z is not None:
        _chk(z.dtype == u.dtype and z.is_cuda and tuple(z.shape) == (batch, dim, L) and (z.stride(-1) == 1 or L == 1),
             "selective_scan: z must match u")
    return batch, dim, L, N, G


def _fill_fwd(P: _lib.ScanFwdParams, u, delta, A, B, C, D, delta_bias, z, delta_softplus, out_dtype, dims):
    batch, dim, L, N, G = dims
    P.batch, P.dim, P.seqlen, P.dstate, P.ngroups = batch, dim, L, N, G
    P.in_dtype, P.out_dtype, P.delta_softplus = _DT[u.dtype], _DT[out_dtype], int(bool(delta_softplus))
    P.u, P.delta, P.A, P.B, P.C = u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr()
    P.D, P.delta_bias, P.z = _ptr(D), _ptr(delta_bias), _ptr(z)
    P.u_bstride, P.u_dstride = u.stride(0), u.stride(1)
    P.delta_bstride, P.delta_dstride = delta.stride(0), delta.stride(1)
    P.B_bstride, P.B_gstride, P.B_nstride = B.stride(0), B.stride(1), B.stride(2)
    P.C_bstride, P.C_gstride, P.C_nstride = C.stride(0), C.stride(1), C.stride(2)
    if z is not None:
        P.z_bstride, P.z_dstride = z.stride(0), z.stride(1)


def _n_ref(L):
    return (L + _lib.REF_CHUNK - 1) // _lib.REF_CHUNK


def _n_ckpt(batch, dim, L, N):
    """fp32 elements of the library's checkpoint workspace for this shape (layout private to the library)."""
    return int(_lib.lib().ss2d_scan_ckpt_floats(batch, dim, L, N))


def _alloc_x_ckpt(u, batch, dim, L, N, family):
    """One fp32 buffer = [ x (batch,dim,n_ref,2N) | ckpt workspace | `family` pad words ].

    ``x`` is the contiguous leading view, so callers that only know the reference contract
    (``last_state = x[:, :, -1, 1::2]``, test_selective_scan.py:79) see exactly the reference tensor, while
    ``scan_bwd`` can find the checkpoints behind it (``_ckpt_of``) even when ``x`` travelled through the
    reference's own autograd Function (ITS/models/vmamba_layers.py:184,190).  The number of pad words (1 or 2) records
    which kernel family wrote the checkpoints — their layouts differ, and the backward must read them the way the
    forward wrote them whatever the dispatch rule says at that time."""
    nx, nc = batch * dim * _n_ref(L) * 2 * N, _n_ckpt(batch, dim, L, N)
    buf = torch.empty(nx + nc + family, device=u.device, dtype=torch.float32)
    x = buf[:nx].view(batch, dim, _n_ref(L), 2 * N)
    ckpt = buf[nx:nx + nc]
    return x, ckpt


def _ckpt_of(x: Optional[torch.Tensor], batch, dim, L, N):
    """-> (ckpt view, family) behind an ``x`` produced by scan_fwd, or (None, 0) for a foreign tensor."""
    if x is None or x.dtype != torch.float32 or not x.is_cuda:
        return None, 0
    nx, nc = batch * dim * _n_ref(L) * 2 * N, _n_ckpt(batch, dim, L, N)
    if x.storage_offset() != 0 or x.numel() != nx or not x.is_contiguous():
        return None, 0
    # the storage size (x | ckpt | 1 or 2 pad words) is the signature; no device read, hence no host sync
    family = x.untyped_storage().nbytes() // 4 - (nx + nc)
    if x.untyped_storage().nbytes() % 4 or family not in (_lib.FAMILY_STATELANES, _lib.FAMILY_WARPSCAN):
        return None, 0
    flat = torch.as_strided(x, (nx + nc + family,), (1,), 0)
    return flat[nx:nx + nc], family


class ScanCkpt:
    """The forward's checkpoint workspace and the kernel family that wrote it (what scan_bwd needs back)."""
    __slots__ = ("data", "family")

    def __init__(self, data, family):
        self.data, self.family = data, family


def scan_fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, out_float=True, z=None, family=0):
    """-> (out, x, ckpt, out_z).  `fwd` of selective_scan_cuda_oflex (selective_scan_oflex.cpp:157-243) plus the
    z gate of the API of record.  `nrows` is accepted and ignored like the reference (:236-238).  `ckpt` is a ScanCkpt;
    `family` (tests) pins the kernel family, 0 lets the library choose by problem size."""
    dims = _check_scan_inputs(u, delta, A, B, C, D, delta_bias, z)
    batch, dim, L, N, G = dims
    out_dtype = torch.float32 if out_float else u.dtype
    out = torch.empty((batch, dim, L), device=u.device, dtype=out_dtype)
    out_z = torch.empty_like(out) if z is not None else None
    P = _lib.ScanFwdParams()
    _fill_fwd(P, u, delta, A, B, C, D, delta_bias, z, delta_softplus, out_dtype, dims)
    P.family = family
    P.family = family = int(_lib.lib().ss2d_scan_family(C_byref(P)))  # resolved once; the backward gets the same value
    x, ckpt = _alloc_x_ckpt(u, batch, dim, L, N, family)
    P.out, P.out_bstride, P.out_dstride = out.data_ptr(), out.stride(0), out.stride(1)
    P.out_z, P.x, P.ckpt = _ptr(out_z), x.data_ptr(), ckpt.data_ptr()
    with torch.cuda.device(u.device):
        _lib.check(_lib.lib().ss2d_selective_scan_fwd(C_byref(P), _stream(u)), "ss2d_selective_scan_fwd")
    return out, x, ScanCkpt(ckpt, family), out_z


def C_byref(s):
    return C.cast(C.pointer(s), C.c_void_p)


def scan_bwd(u, delta, A, B, C, D, delta_bias, dout, x=None, delta_softplus=False, nrows=1, ckpt=None, z=None, out=None):
    """-> (du, ddelta, dA, dB, dC, dD, ddelta_bias, dz).  `bwd` of selective_scan_cuda_oflex
    (selective_scan_oflex.cpp:245-358): du/ddelta in the input dtype, dA/dD/ddelta_bias fp32, dB/dC cast to the
    dtype of B/C (:356).  `dout` may be fp32 or the input dtype (:260)."""
    dims = _check_scan_inputs(u, delta, A, B, C, D, delta_bias, z)
    batch, dim, L, N, G = dims
    _chk(dout.is_cuda and (dout.dtype == u.dtype or dout.dtype == torch.float32), "selective_scan: bad dout dtype")
    _chk(tuple(dout.shape) == (batch, dim, L), "selective_scan: dout must have shape (batch, dim, seqlen)")
    _chk(dout.stride(-1) == 1 or L == 1, "selective_scan: dout.stride(-1) must be 1")
    if _n_ref(L) > 1 and ckpt is None:
        _chk(x is not None, "selective_scan: x is required when seqlen > 2048")  # oflex.cpp:315
    if x is not None:
        _chk(x.dtype == torch.float32 and x.is_cuda and tuple(x.shape) == (batch, dim, _n_ref(L), 2 * N),
             "selective_scan: x must be float32 (batch, dim, n_chunks, 2*dstate)")
    if z is not None:
        _chk(out is not None and out.dtype == dout.dtype, "selective_scan: the un-gated out is required with z")
    if isinstance(ckpt, ScanCkpt):
        ckpt, family = ckpt.data, ckpt.family
    else:
        ckpt, family = _ckpt_of(x, batch, dim, L, N)  # (None, 0) for a foreign x: the library rebuilds the checkpoints
    scratch = None
    if ckpt is None and L > _lib.SL_BLOCK:
        scratch = torch.empty(_n_ckpt(batch, dim, L, N), device=u.device, dtype=torch.float32)
    du, ddelta = torch.empty_like(u, memory_format=torch.contiguous_format), torch.empty_like(delta, memory_format=torch.contiguous_format)
    # the accumulated gradients share ONE zero-filled fp32 buffer (one memset instead of five)
    n_bc, n_a = batch * G * N * L, dim * N
    acc = torch.zeros(2 * n_bc + n_a + 2 * dim, device=u.device, dtype=torch.float32)
    dB, dC = acc[:n_bc].view(batch, G, N, L), acc[n_bc:2 * n_bc].view(batch, G, N, L)
    dA = acc[2 * n_bc:2 * n_bc + n_a].view(dim, N)
    dD = acc[2 * n_bc + n_a:2 * n_bc + n_a + dim] if D is not None else None
    dbias = acc[2 * n_bc + n_a + dim:] if delta_bias is not None else None
    dz = torch.empty_like(u, memory_format=torch.contiguous_format) if z is not None else None
    P = _lib.ScanBwdParams()
    _fill_fwd(P.f, u, delta, A, B, C, D, delta_bias, z, delta_softplus, dout.dtype, dims)
    if out is not None:
        P.f.out, P.f.out_bstride, P.f.out_dstride = out.data_ptr(), out.stride(0), out.stride(1)
    P.f.ckpt, P.f.family = _ptr(ckpt), family
    P.dout, P.dout_bstride, P.dout_dstride = dout.data_ptr(), dout.stride(0), dout.stride(1)
    P.ckpt_scratch = _ptr(scratch)
    P.du, P.ddelta, P.dz = du.data_ptr(), ddelta.data_ptr(), _ptr(dz)
    P.dA, P.dB, P.dC, P.dD, P.ddelta_bias = dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), _ptr(dD), _ptr(dbias)
    with torch.cuda.device(u.device):
        _lib.check(_lib.lib().ss2d_selective_scan_bwd(C_byref(P), _stream(u)), "ss2d_selective_scan_bwd")
    return du, ddelta, dA, dB.to(B.dtype), dC.to(C.dtype), dD, dbias, dz


def build_selective_scan_fn(mode: str = "ssoflex", out_float: bool = True, tag=None):
    """Same factory shape as the reference's build_selective_scan_fn (test_selective_scan.py:18-165):
    mode "ssoflex" -> fp32 `out` from the kernel then cast back to u.dtype (:158-159); "sscore" -> out in u.dtype."""
    ssoflex = mode == "ssoflex"

    class SelectiveScanFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                    return_last_state=False, nrows=1, backnrows=-1):
            # the kernels need unit stride along L only (selective_scan_oflex.cpp:181-182,198-200)
            u, delta, B, C, z = (t if t is None or t.stride(-1) == 1 else t.contiguous() for t in (u, delta, B, C, z))
            D = None if D is None else D.contiguous()
            ctx.squeeze_B = B.dim() == 3
            ctx.squeeze_C = C.dim() == 3
            if ctx.squeeze_B:
                B = B.unsqueeze(1)
            if ctx.squeeze_C:
                C = C.unsqueeze(1)
            ctx.d_dtype = D.dtype if D is not None else None
            ctx.bias_dtype = delta_bias.dtype if delta_bias is not None else None
            if D is not None and D.dtype != torch.float32:
                D = D.float()
            if delta_bias is not None and delta_bias.dtype != torch.float32:
                delta_bias = delta_bias.float()
            assert u.shape[1] % (B.shape[1] * nrows) == 0
            assert nrows in [1, 2, 3, 4]
            out, x, ckpt, out_z = scan_fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows,
                                           out_float=(ssoflex and out_float), z=z)
            ctx.delta_softplus = delta_softplus
            ctx.has_z = z is not None
            last_state = x[:, :, -1, 1::2]  # (batch, dim, dstate)
            if ctx.has_z:
                ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, ckpt.data, z, out)
                res = out_z
            else:
                ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, ckpt.data)
                res = out
            ctx.family = ckpt.family
            if return_last_state:
                ctx.mark_non_differentiable(last_state)
                return res, last_state
            return res

        @staticmethod
        def backward(ctx, dout, *args):
            if ctx.has_z:
                u, delta, A, B, C, D, delta_bias, ckpt, z, out = ctx.saved_tensors
            else:
                u, delta, A, B, C, D, delta_bias, ckpt = ctx.saved_tensors
                z = out = None
            if dout.stride(-1) != 1:
                dout = dout.contiguous()
            du, ddelta, dA, dB, dC, dD, dbias, dz = scan_bwd(u, delta, A, B, C, D, delta_bias, dout, None, ctx.delta_softplus, 1,
                                                             ckpt=ScanCkpt(ckpt, ctx.family), z=z, out=out)
            if ctx.squeeze_B:
                dB = dB.squeeze(1)
            if ctx.squeeze_C:
                dC = dC.squeeze(1)
            if dD is not None and ctx.d_dtype != dD.dtype:
                dD = dD.to(ctx.d_dtype)
            if dbias is not None and ctx.bias_dtype != dbias.dtype:
                dbias = dbias.to(ctx.bias_dtype)
            return du, ddelta, dA, dB, dC, dD, dz, dbias, None, None, None, None

    def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                          return_last_state=False, nrows=1, backnrows=-1):
        """if return_last_state is True, returns (out, last_state); last_state has shape (batch, dim, dstate)
        and its gradient is not considered in the backward pass (test_selective_scan.py:152-161)."""
        outs = SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state, nrows,
                                     backnrows)
        if ssoflex:
            return outs.to(u.dtype) if not return_last_state else (outs[0].to(u.dtype), outs[1])
        return outs

    selective_scan_fn.__repr__ = lambda *_: f"selective_scan_fn | {mode} | {tag}"
    return selective_scan_fn


selective_scan_fn = build_selective_scan_fn("ssoflex")
