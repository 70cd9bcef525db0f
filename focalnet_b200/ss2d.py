"""Host-side mirror of the SS2D core seams S2 and S3 (ITS/models/vmamba_layers.py, ITS/models/csm_triton.py).

* ``CrossScan`` / ``CrossMerge`` (aliases ``CrossScanTriton`` / ``CrossMergeTriton``): autograd Functions with the
  reference's signatures (csm_triton.py:163-210; torch twins vmamba_layers.py:29-71) on the CUDA kernels of
  ``csrc/ss2d_cross.cu`` — no Triton.
* ``cross_selective_scan``: same signature and result as the reference function (vmamba_layers.py:200-299) but
  the 4-direction CrossScan / CrossMerge are folded into the scan kernels' addressing
  (``ss2d_cross_scan_fwd/_bwd``): the (B,4,D,L) copies of x and of the scan outputs are never materialised.
* ``dwconv_silu``: the permute + depthwise 3x3 conv + SiLU pre-mix of SS2D.forwardv2 (vmamba_layers.py:591-594).
* ``patch_ss2d``: re-binds ``forward_core`` of every SS2D in an *unchanged* reference model to the fused path.

torch is plumbing (allocation, the two skinny projection GEMMs through cuBLAS, LayerNorm); no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from functools import partial

import torch
import torch.nn.functional as F

from . import _lib
from .selective_scan import _DT, C_byref, ScanCkpt, _chk, _n_ckpt, _ptr, _stream, build_selective_scan_fn, scan_bwd, scan_fwd


# --------------------------------------------------------------------------------------------- S2
def _cross(op: str, src: torch.Tensor, dst: torch.Tensor, B, Cn, H, W):
    _chk(src.is_cuda and src.dtype in _DT, "cross scan/merge: CUDA float32/float16/bfloat16 tensors only (no CPU path)")
    fn = getattr(_lib.lib(), op)
    with torch.cuda.device(src.device):
        _lib.check(fn(src.data_ptr(), dst.data_ptr(), B, Cn, H, W, _DT[src.dtype], _stream(src)), op)
    return dst


def cross_scan(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) -> (B,4,C,H*W)"""
    B, Cn, H, W = x.shape
    x = x.contiguous()
    return _cross("ss2d_cross_scan", x, x.new_empty((B, 4, Cn, H * W)), B, Cn, H, W)


def cross_merge(ys: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """(B,4,C,H*W) -> (B,C,H*W)"""
    B, K, Cn = ys.shape[:3]
    _chk(K == 4, "cross_merge expects 4 directions")
    ys = ys.contiguous()
    return _cross("ss2d_cross_merge", ys, ys.new_empty((B, Cn, H * W)), B, Cn, H, W)


class CrossScan(torch.autograd.Function):
    """CrossScanTriton (csm_triton.py:163-185): forward (B,C,H,W)->(B,4,C,L); backward is a cross-merge."""

    @staticmethod
    def forward(ctx, x: torch.Tensor):
        ctx.shape = x.shape
        return cross_scan(x)

    @staticmethod
    def backward(ctx, ys: torch.Tensor):
        B, Cn, H, W = ctx.shape
        return cross_merge(ys, H, W).view(B, Cn, H, W)


class CrossMerge(torch.autograd.Function):
    """CrossMergeTriton (csm_triton.py:188-210): forward (B,4,C,H,W)->(B,C,L); backward is a cross-scan."""

    @staticmethod
    def forward(ctx, ys: torch.Tensor):
        B, K, Cn, H, W = ys.shape
        ctx.shape = (B, Cn, H, W)
        return cross_merge(ys.reshape(B, K, Cn, H * W), H, W)

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        B, Cn, H, W = ctx.shape
        return cross_scan(dy.reshape(B, Cn, H, W)).view(B, 4, Cn, H, W)


CrossScanTriton, CrossMergeTriton = CrossScan, CrossMerge


# ------------------------------------------------------------------------------------- selective scan seams
class _SelectiveScanBase(torch.autograd.Function):
    """SelectiveScanOflex / SelectiveScanCore (vmamba_layers.py:155-196) on the C ABI."""
    OFLEX = True

    @classmethod
    def _fwd(cls, ctx, u, delta, A, B, Cm, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
        ctx.delta_softplus = delta_softplus
        out, x, ckpt, _ = scan_fwd(u, delta, A, B, Cm, D, delta_bias, delta_softplus, 1, out_float=(oflex and cls.OFLEX))
        ctx.save_for_backward(u, delta, A, B, Cm, D, delta_bias, ckpt.data)
        ctx.family = ckpt.family
        return out

    @staticmethod
    def _bwd(ctx, dout):
        u, delta, A, B, Cm, D, delta_bias, ckpt = ctx.saved_tensors
        if dout.stride(-1) != 1:
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, dbias, _ = scan_bwd(u, delta, A, B, Cm, D, delta_bias, dout, None, ctx.delta_softplus, 1,
                                                        ckpt=ScanCkpt(ckpt, ctx.family))
        return du, ddelta, dA, dB, dC, dD, dbias, None, None, None, None


class SelectiveScanOflex(_SelectiveScanBase):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
        return SelectiveScanOflex._fwd(ctx, u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, backnrows, oflex)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout, *args):
        return _SelectiveScanBase._bwd(ctx, dout)


class SelectiveScanCore(_SelectiveScanBase):
    OFLEX = False

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
        return SelectiveScanCore._fwd(ctx, u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, backnrows, oflex)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout, *args):
        return _SelectiveScanBase._bwd(ctx, dout)


# --------------------------------------------------------------------------------------------- S3
class FusedCrossScanFn(torch.autograd.Function):
    """x (B,D,H,W) spatial, delta (B,4D,L) / Bs, Cs (B,4,N,L) in scan order  ->  y (B,D,L) spatial fp32 =
    CrossMerge(selective_scan(CrossScan(x), ...)) without the 4x copies (ss2d_cross_scan_fwd/_bwd).

    ``FusedCrossScanFn.recompute = True`` trades one extra states-only forward sweep in the backward for not keeping the
    block-boundary states between the passes (activation memory of the scan drops to the inputs autograd holds anyway)."""
    recompute = False

    @staticmethod
    def forward(ctx, x, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus=True, deterministic=False):
        B, D, H, W = x.shape
        L, N = H * W, A.shape[1]
        _chk(x.is_cuda and x.dtype in _DT, "fused SS2D core: x must be a CUDA float32/float16/bfloat16 tensor")
        _chk(delta.dtype == x.dtype and Bs.dtype == x.dtype and Cs.dtype == x.dtype, "fused SS2D core: dtype mismatch")
        _chk(tuple(delta.shape) == (B, 4 * D, L) and tuple(A.shape) == (4 * D, N), "fused SS2D core: bad delta / A shape")
        _chk(tuple(Bs.shape) == (B, 4, N, L) and tuple(Cs.shape) == (B, 4, N, L), "fused SS2D core: bad B / C shape")
        x, delta, A = x.contiguous(), delta.contiguous(), A.contiguous().float()
        if not (Bs.stride(3) == 1 and Bs.stride(2) == L and Bs.stride() == Cs.stride()):
            Bs, Cs = Bs.contiguous(), Cs.contiguous()
        Ds = None if Ds is None else Ds.contiguous().float()
        delta_bias = None if delta_bias is None else delta_bias.contiguous().float()
        y = torch.zeros((B, D, L), device=x.device, dtype=torch.float32)
        # block-boundary states for the backward: not written at all when nothing needs a gradient (inference), nor in
        # `recompute` mode, where the backward rebuilds them with a states-only sweep instead of keeping them alive
        # between the two passes (as many bytes as delta: 1.6 GB per call at B=32, L=16384)
        keep = any(ctx.needs_input_grad) and not FusedCrossScanFn.recompute
        ckpt = torch.empty(_n_ckpt(B, 4 * D, L, N), device=x.device, dtype=torch.float32) if keep else None
        P = _lib.CrossFwdParams()
        FusedCrossScanFn._fill(P, x, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus, (B, D, H, W, N))
        P.y, P.ckpt = y.data_ptr(), _ptr(ckpt)
        # the family is resolved ONCE here and handed to the backward (the checkpoint layouts of the families differ)
        P.family = family = int(_lib.lib().ss2d_cross_family(B, D, H, W, N, _DT[x.dtype], 0))
        P.deterministic = int(bool(deterministic))
        work = FusedCrossScanFn._work(x, N, L, False, family)
        P.work = _ptr(work)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ss2d_cross_scan_fwd(C_byref(P), _stream(x)), "ss2d_cross_scan_fwd")
        ctx.save_for_backward(x, delta, A, Bs, Cs, Ds, delta_bias, ckpt)
        ctx.delta_softplus, ctx.family, ctx.deterministic = delta_softplus, family, bool(deterministic)
        return y

    @staticmethod
    def _work(x, N, L, backward, family):
        """Scratch for the state-lanes kernels (x^T | y^T, resp. x^T | dy^T | dx^T); None with the warp-scan kernels."""
        B, D, H, W = x.shape
        n = int(_lib.lib().ss2d_cross_work_floats(B, D, H, W, N, _DT[x.dtype], int(backward), family))
        return torch.empty(n, device=x.device, dtype=torch.float32) if n else None

    @staticmethod
    def _fill(P, x, delta, A, Bs, Cs, Ds, delta_bias, delta_softplus, dims):
        B, D, H, W, N = dims
        P.batch, P.D, P.H, P.W, P.dstate = B, D, H, W, N
        P.in_dtype, P.delta_softplus = _DT[x.dtype], int(bool(delta_softplus))
        P.x, P.delta, P.B, P.C, P.A = x.data_ptr(), delta.data_ptr(), Bs.data_ptr(), Cs.data_ptr(), A.data_ptr()
        P.Dskip, P.delta_bias = _ptr(Ds), _ptr(delta_bias)
        P.bc_bstride, P.bc_gstride = Bs.stride(0), Bs.stride(1)

    @staticmethod
    def backward(ctx, dy):
        x, delta, A, Bs, Cs, Ds, delta_bias, ckpt = ctx.saved_tensors
        B, D, H, W = x.shape
        L, N = H * W, A.shape[1]
        dy = dy.contiguous().float()
        if ckpt is None:  # recompute mode: states-only forward sweep (y = NULL) into a transient buffer
            ckpt = torch.empty(_n_ckpt(B, 4 * D, L, N), device=x.device, dtype=torch.float32)
            Pf = _lib.CrossFwdParams()
            FusedCrossScanFn._fill(Pf, x, delta, A, Bs, Cs, Ds, delta_bias, ctx.delta_softplus, (B, D, H, W, N))
            Pf.y, Pf.ckpt, Pf.family = None, ckpt.data_ptr(), ctx.family
            wf = FusedCrossScanFn._work(x, N, L, False, ctx.family)
            Pf.work = _ptr(wf)
            with torch.cuda.device(x.device):
                _lib.check(_lib.lib().ss2d_cross_scan_fwd(C_byref(Pf), _stream(x)), "ss2d_cross_scan_fwd (states only)")
        # every accumulator the kernels add into comes out of ONE zero-filled buffer (one fill launch instead of six);
        # pieces start on 16-byte boundaries (the kernels use 128-bit reductions)
        sizes = [B * D * L, A.numel(), B * 4 * N * L, B * 4 * N * L, 4 * D if Ds is not None else 0, 4 * D if delta_bias is not None else 0]
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot)
            tot += (n + 3) // 4 * 4
        acc = torch.zeros(tot, device=x.device, dtype=torch.float32)
        dx, dA, dB, dC, dDs, dbias = (acc[o:o + n] if n else None for o, n in zip(offs, sizes))
        dx, dA, dB, dC = dx.view(B, D, L), dA.view_as(A), dB.view(B, 4, N, L), dC.view(B, 4, N, L)
        ddelta = torch.empty_like(delta)
        P = _lib.CrossBwdParams()
        FusedCrossScanFn._fill(P.f, x, delta, A, Bs, Cs, Ds, delta_bias, ctx.delta_softplus, (B, D, H, W, N))
        P.f.ckpt, P.f.family, P.f.deterministic = ckpt.data_ptr(), ctx.family, int(ctx.deterministic)
        work = FusedCrossScanFn._work(x, N, L, True, ctx.family)
        P.f.work = _ptr(work)
        P.dy, P.dx, P.ddelta = dy.data_ptr(), dx.data_ptr(), ddelta.data_ptr()
        P.dA, P.dB, P.dC, P.dDskip, P.ddelta_bias = dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), _ptr(dDs), _ptr(dbias)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().ss2d_cross_scan_bwd(C_byref(P), _stream(x)), "ss2d_cross_scan_bwd")
        return dx.view(B, D, H, W).to(x.dtype), ddelta, dA, dB.to(Bs.dtype), dC.to(Cs.dtype), dDs, dbias, None, None


class MergeNormGateFn(torch.autograd.Function):
    """y (B,D,L) fp32 spatial  ->  LayerNorm_D(y^T) [* silu(z)]  as (B,L,D) channels-last, one kernel
    (ss2d_merge_norm_gate_fwd/_bwd).  `z` is a channels-last view whose last dim has stride 1 (e.g. the z-half
    ``xz[..., D:]`` of the in_proj output); its pixel stride is passed through, no copy is made."""

    @staticmethod
    def forward(ctx, y, weight, bias, eps, z):
        B, D, L = y.shape
        _chk(y.is_cuda and y.dtype == torch.float32, "merge_norm_gate: y must be a float32 CUDA tensor (no CPU path)")
        y = y.contiguous()
        w, bb = weight.contiguous().float(), bias.contiguous().float()
        zs, zshape = 0, None
        if z is not None:
            zshape = tuple(z.shape)
            _chk(z.dtype == torch.float32 and z.is_cuda and z.shape[-1] == D and z.numel() == B * L * D and z.stride(-1) == 1,
                 "merge_norm_gate: z must be float32 channels-last with D channels per pixel")
            zz = z.reshape(B * L, D) if z.dim() != 2 else z
            if zz.stride(1) != 1 or (B * L > 1 and zz.stride(0) < D):
                zz = zz.contiguous()
            z, zs = zz, zz.stride(0) if B * L > 1 else D
        out = torch.empty((B, L, D), device=y.device, dtype=torch.float32)
        with torch.cuda.device(y.device):
            _lib.check(_lib.lib().ss2d_merge_norm_gate_fwd(y.data_ptr(), w.data_ptr(), bb.data_ptr(), float(eps), _ptr(z), zs,
                                                          out.data_ptr(), B, D, L, _stream(y)), "ss2d_merge_norm_gate_fwd")
        ctx.save_for_backward(y, w, bb, z)
        ctx.eps, ctx.zs, ctx.zshape = float(eps), zs, zshape
        return out

    @staticmethod
    def backward(ctx, dout):
        y, w, bb, z = ctx.saved_tensors
        B, D, L = y.shape
        dout = dout.contiguous().float()
        dy = torch.empty_like(y)
        dwb = torch.zeros(2 * D, device=y.device, dtype=torch.float32)
        dz = torch.empty((B * L, D), device=y.device, dtype=torch.float32) if z is not None else None
        with torch.cuda.device(y.device):
            _lib.check(_lib.lib().ss2d_merge_norm_gate_bwd(y.data_ptr(), w.data_ptr(), bb.data_ptr(), ctx.eps, _ptr(z), ctx.zs,
                                                          dout.data_ptr(), dy.data_ptr(), _ptr(dz), D, dwb.data_ptr(),
                                                          dwb[D:].data_ptr(), B, D, L, _stream(y)), "ss2d_merge_norm_gate_bwd")
        return dy, dwb[:D], dwb[D:], None, (dz.view(ctx.zshape) if dz is not None else None)


def merge_norm_gate(y, weight, bias, eps=1e-5, z=None):
    """LayerNorm over channels of the merged scan output y (B,D,L), transposed to (B,L,D), optionally * silu(z)."""
    B, D, L = y.shape
    out = MergeNormGateFn.apply(y, weight, bias, eps, z)
    return out


class ToScanOrderFn(torch.autograd.Function):
    """t (B,4,C,L), plane k in SPATIAL order -> each direction's own scan order (ss2d_cross_permute); the backward is the
    inverse permutation.  Only ever applied to the small x_dbl (R+2N = 38 rows per direction), never to x (192 rows)."""

    @staticmethod
    def forward(ctx, t, H, W):
        B, K, Cn, L = t.shape
        _chk(K == 4 and L == H * W and t.is_cuda and t.dtype in _DT, "to_scan_order: (B,4,C,H*W) CUDA tensor expected")
        t = t.contiguous()
        out = torch.empty_like(t)
        ctx.dims = (B, Cn, H, W)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().ss2d_cross_permute(t.data_ptr(), out.data_ptr(), B, Cn, H, W, _DT[t.dtype], 0, _stream(t)),
                       "ss2d_cross_permute")
        return out

    @staticmethod
    def backward(ctx, g):
        B, Cn, H, W = ctx.dims
        g = g.contiguous()
        out = torch.empty_like(g)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().ss2d_cross_permute(g.data_ptr(), out.data_ptr(), B, Cn, H, W, _DT[g.dtype], 1, _stream(g)),
                       "ss2d_cross_permute")
        return out, None, None


def _to_scan_order(t: torch.Tensor, H: int, W: int) -> torch.Tensor:
    return ToScanOrderFn.apply(t, H, W)


class DtProjFn(torch.autograd.Function):
    """dts_lr (B,K,R,L) fp32 (any batch / direction / rank strides, unit stride along L), W (K,D,R)  ->  (B, K*D, L):
    the low-rank dt projection ``F.conv1d(dts.view(B, K*R, L), W.view(K*D, R, 1), groups=K)`` of vmamba_layers.py:264 as one
    HBM stream each way (ss2d_dt_proj_fwd/_bwd) instead of a grouped cuDNN convolution with layout conversions."""

    @staticmethod
    def supported(dts_lr, W):
        B, K, R, L = dts_lr.shape
        return (dts_lr.is_cuda and dts_lr.dtype == torch.float32 and W.dtype == torch.float32 and not torch.is_autocast_enabled()
                and R <= 8 and W.shape[1] <= 512 and L % 4 == 0 and dts_lr.stride(3) == 1 and dts_lr.data_ptr() % 16 == 0
                and all(st % 4 == 0 for st in dts_lr.stride()[:3]))

    @staticmethod
    def forward(ctx, dts_lr, W):
        B, K, R, L = dts_lr.shape
        D = W.shape[1]
        Wc = W.contiguous()
        out = torch.empty((B, K * D, L), device=dts_lr.device, dtype=torch.float32)
        with torch.cuda.device(dts_lr.device):
            _lib.check(_lib.lib().ss2d_dt_proj_fwd(dts_lr.data_ptr(), dts_lr.stride(0), dts_lr.stride(1), dts_lr.stride(2), Wc.data_ptr(),
                                                   out.data_ptr(), B, K, D, R, L, _stream(dts_lr)), "ss2d_dt_proj_fwd")
        ctx.save_for_backward(dts_lr, Wc)
        return out

    @staticmethod
    def backward(ctx, dout):
        dts_lr, Wc = ctx.saved_tensors
        B, K, R, L = dts_lr.shape
        D = Wc.shape[1]
        dout = dout.contiguous()
        dx = torch.empty((B, K, R, L), device=dout.device, dtype=torch.float32)
        dW = torch.zeros_like(Wc)
        with torch.cuda.device(dout.device):
            _lib.check(_lib.lib().ss2d_dt_proj_bwd(dout.data_ptr(), dts_lr.data_ptr(), dts_lr.stride(0), dts_lr.stride(1), dts_lr.stride(2),
                                                   Wc.data_ptr(), dx.data_ptr(), dW.data_ptr(), B, K, D, R, L, _stream(dout)),
                       "ss2d_dt_proj_bwd")
        return dx, dW


def _dt_proj(dts_lr, dt_projs_weight, no_einsum=True):
    """(B,K,R,L) x (K,D,R) -> (B, K*D, L).  fp32 outside autocast: this library's stream kernels (exact fp32 — the library
    conv1d would round to TF32 under torch's default cuDNN policy); otherwise the reference's own library call."""
    B, K, R, L = dts_lr.shape
    if DtProjFn.supported(dts_lr, dt_projs_weight):
        return DtProjFn.apply(dts_lr, dt_projs_weight)
    if no_einsum:
        return F.conv1d(dts_lr.reshape(B, K * R, L), dt_projs_weight.reshape(-1, R, 1), groups=K)
    return torch.einsum("b k r l, k d r -> b k d l", dts_lr, dt_projs_weight).reshape(B, -1, L)


def cross_selective_scan(
    x: torch.Tensor = None,
    x_proj_weight: torch.Tensor = None,
    x_proj_bias: torch.Tensor = None,
    dt_projs_weight: torch.Tensor = None,
    dt_projs_bias: torch.Tensor = None,
    A_logs: torch.Tensor = None,
    Ds: torch.Tensor = None,
    delta_softplus=True,
    out_norm: torch.nn.Module = None,
    out_norm_shape="v0",
    to_dtype=True,
    force_fp32=False,
    nrows=-1,
    backnrows=-1,
    ssoflex=True,
    SelectiveScan=None,
    CrossScan=None,
    CrossMerge=None,
    no_einsum=False,
    dt_low_rank=True,
):
    """Drop-in for the reference ``cross_selective_scan`` (vmamba_layers.py:200-299): same arguments, same
    (B,H,W,D) result.  ``SelectiveScan`` / ``CrossScan`` / ``CrossMerge`` / ``nrows`` / ``backnrows`` select
    implementations in the reference; here they are accepted and ignored — there is one implementation.  The flags
    that change NUMERICS are honoured like the reference: ``force_fp32`` (cast x / dts / Bs / Cs to float after the
    projections, :281-285), ``no_einsum`` (projections through 1x1 conv1d = cuDNN and its TF32 policy, else through
    einsum = cuBLAS and its policy, :260-270), ``ssoflex`` (fp32 scan output, else the input dtype), ``to_dtype``.

    x_proj is pointwise over pixels, so x_dbl_k = W_k x is computed ONCE from x in spatial order and only that
    38-row tensor is permuted into each direction's scan order; x itself (192 rows) is read by the scan kernel
    with the direction's addressing."""
    B, D, H, W = x.shape
    N = A_logs.shape[1]
    K, _, R = dt_projs_weight.shape
    L = H * W
    _chk(K == 4 and dt_low_rank, "fused SS2D core supports the 4-direction low-rank-dt configuration of the model")
    xf = x.reshape(B, D, L)
    C_all = K * (R + 2 * N)
    # the two projections stay on the reference's own library calls (vmamba_layers.py:262-270), run in x's dtype like
    # there — x_proj un-grouped, because all four directions read the same spatial-order x
    if no_einsum:
        x_dbl = F.conv1d(xf, x_proj_weight.reshape(C_all, D, 1), None if x_proj_bias is None else x_proj_bias.reshape(-1))
    else:
        x_dbl = torch.einsum("b d l, c d -> b c l", xf, x_proj_weight.reshape(C_all, D))
        if x_proj_bias is not None:
            x_dbl = x_dbl + x_proj_bias.reshape(1, -1, 1)
    x_dbl = _to_scan_order(x_dbl.view(B, K, R + 2 * N, L), H, W)        # (B, K, R+2N, L), each direction's scan order
    dts_lr, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = _dt_proj(dts_lr, dt_projs_weight, no_einsum)
    As = -torch.exp(A_logs.to(torch.float))
    xin = x
    if force_fp32:  # vmamba_layers.py:281-285
        xin, dts, Bs, Cs = xin.to(torch.float), dts.to(torch.float), Bs.to(torch.float), Cs.to(torch.float)
    y = FusedCrossScanFn.apply(xin, dts, As, Bs, Cs, Ds.to(torch.float), dt_projs_bias.reshape(-1).to(torch.float),
                               delta_softplus)
    if not ssoflex:
        y = y.to(xin.dtype)  # SSOflex with oflex=False / SSCore return the scan output in the input dtype
    if out_norm_shape in ["v1"]:
        y = out_norm(y.view(B, -1, H, W)).permute(0, 2, 3, 1)
    elif y.dtype == torch.float32 and isinstance(out_norm, torch.nn.LayerNorm) and out_norm.elementwise_affine and \
            out_norm.bias is not None and out_norm.weight.dtype == torch.float32 and tuple(out_norm.normalized_shape) == (D,) \
            and D <= 512 and not torch.is_autocast_enabled():
        # transpose + LayerNorm in one kernel (row N1 of SURVEY §8f) instead of a transpose copy + ATen LayerNorm
        y = merge_norm_gate(y, out_norm.weight, out_norm.bias, out_norm.eps).view(B, H, W, -1)
    else:
        y = out_norm(y.transpose(1, 2).contiguous()).view(B, H, W, -1)
    return y.to(x.dtype) if to_dtype else y


def block_supported(m: torch.nn.Module) -> bool:
    """True when ``ss2d_forward`` computes exactly what this (unchanged reference) SS2D module's forwardv2 does: the
    shipped ITS configuration — forward type v4 (force_fp32=False, no_einsum, ssoflex), 3x3 depthwise conv with bias,
    SiLU on x and z, z gate, LayerNorm(d_inner) in channels-last ("v0"), fp32 parameters, no dropout."""
    try:
        fc = m.forward_core
        kw = getattr(fc, "keywords", {}) or {}
        conv = m.conv2d
        return (
            m.d_conv == 3 and not m.disable_z and not m.disable_z_act and isinstance(m.act, torch.nn.SiLU)
            and m.out_norm_shape == "v0" and isinstance(m.out_norm, torch.nn.LayerNorm) and m.out_norm.elementwise_affine
            and m.out_norm.bias is not None and tuple(m.out_norm.normalized_shape) == (conv.out_channels,)
            and conv.out_channels <= 512 and conv.kernel_size == (3, 3) and conv.padding == (1, 1)
            and conv.groups == conv.in_channels == conv.out_channels and conv.bias is not None
            and isinstance(m.dropout, torch.nn.Identity) and m.in_proj.weight.dtype == torch.float32
            and m.dt_projs_weight.shape[0] == 4 and m.A_logs.shape[1] == 16
            and not kw.get("force_fp32", False) and kw.get("no_einsum", False)
        )
    except AttributeError:
        return False


def ss2d_forward(m: torch.nn.Module, x: torch.Tensor) -> torch.Tensor:
    """SS2D.forwardv2 (vmamba_layers.py:583-601) on this library's kernels, for an *unchanged* reference SS2D module `m`
    in the configuration ``block_supported`` describes: in_proj -> dwconv3x3+SiLU pre-mix (no permute copy) -> fused
    4-direction scan -> transpose+LayerNorm+z-gate epilogue (one kernel) -> out_proj.  Any other configuration (or an
    autocast region, where the reference would run the conv / LayerNorm in 16 bits) takes the module's own forwardv2
    with the fused core."""
    if not block_supported(m) or torch.is_autocast_enabled() or x.dtype != torch.float32:
        return m.forwardv2(x)
    D = m.conv2d.out_channels
    # in_proj as two GEMMs over the two halves of its weight (the same dot products as ONE F.linear + chunk(2), :585-587):
    # the x half and the z half come out as separate dense tensors, so the backward never builds the (B,H,W,2*d_inner)
    # gradient by zero-filling it twice and adding the two halves' contributions (3.5 GB of traffic per call at B=32, 128x128)
    bias = m.in_proj.bias
    xh = F.linear(x, m.in_proj.weight[:D], None if bias is None else bias[:D])     # (B, H, W, d_inner)
    zh = F.linear(x, m.in_proj.weight[D:], None if bias is None else bias[D:])
    B, H, W, _ = xh.shape
    xc = dwconv_silu(xh, m.conv2d.weight, m.conv2d.bias, D)
    N = m.A_logs.shape[1]
    K, _, R = m.dt_projs_weight.shape
    L = H * W
    x_dbl = F.conv1d(xc.reshape(B, D, L), m.x_proj_weight.reshape(K * (R + 2 * N), D, 1))  # the reference's library call
    x_dbl = _to_scan_order(x_dbl.view(B, K, R + 2 * N, L), H, W)
    dts_lr, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = _dt_proj(dts_lr, m.dt_projs_weight)
    y = FusedCrossScanFn.apply(xc, dts, -torch.exp(m.A_logs.float()), Bs, Cs, m.Ds.float(), m.dt_projs_bias.reshape(-1).float(), True)
    y = merge_norm_gate(y, m.out_norm.weight, m.out_norm.bias, m.out_norm.eps, z=zh).view(B, H, W, D)
    return m.dropout(m.out_proj(y))


def patch_ss2d(model: torch.nn.Module, fuse_block: bool = True) -> int:
    """Put every SS2D module of an *unchanged* reference model on this library, without editing ITS/models:

    * ``forward_core`` (an instance attribute, vmamba_layers.py:451) is re-bound to the module's OWN partial with the
      ``cross_selective_scan=`` hook of forward_corev2 (:566) pointing at the fused core — the module's force_fp32 /
      no_einsum / ssoflex choices (:444-449) are kept;
    * with ``fuse_block`` and a module in the shipped configuration (``block_supported``), ``forward`` (also an instance
      attribute, :403) is re-bound to ``ss2d_forward``: the dwconv+SiLU pre-mix and the LayerNorm+gate epilogue run on
      this library's kernels too.

    Returns the number of modules patched; ``unpatch_ss2d`` restores them."""
    n = 0
    for m in model.modules():
        if hasattr(m, "forward_corev2") and callable(getattr(m, "forward_core", None)):
            if not hasattr(m, "_ss2d_b200_orig"):
                m._ss2d_b200_orig = (m.forward_core, m.forward)
            core, fwd = m._ss2d_b200_orig
            m.forward_core = partial(core, cross_selective_scan=cross_selective_scan)
            m.forward = partial(ss2d_forward, m) if (fuse_block and block_supported(m)) else fwd
            n += 1
    return n


def unpatch_ss2d(model: torch.nn.Module) -> int:
    n = 0
    for m in model.modules():
        if hasattr(m, "_ss2d_b200_orig"):
            m.forward_core, m.forward = m._ss2d_b200_orig
            del m._ss2d_b200_orig
            n += 1
    return n


# ------------------------------------------------------------------------------------- dwconv + SiLU pre-mix
class DwConvSiLUFn(torch.autograd.Function):
    """xz (B,H,W,Cs) channels-last (first C channels convolved), weight (C,1,3,3), bias (C)|None
    -> silu(dwconv3x3(x) + bias) as (B,C,H,W)."""

    @staticmethod
    def forward(ctx, xz, weight, bias, Cn):
        _chk(xz.is_cuda and xz.dtype == torch.float32, "dwconv_silu: float32 CUDA tensors only (no CPU path)")
        B, H, W, Cs = xz.shape
        _chk(xz.stride(3) == 1 and xz.is_contiguous(), "dwconv_silu: xz must be contiguous channels-last")
        w = weight.reshape(Cn, 9).contiguous().float()
        bb = None if bias is None else bias.contiguous().float()
        out = torch.empty((B, Cn, H, W), device=xz.device, dtype=torch.float32)
        with torch.cuda.device(xz.device):
            _lib.check(_lib.lib().ss2d_dwconv_silu_fwd(xz.data_ptr(), Cs, w.data_ptr(), _ptr(bb), out.data_ptr(), B, Cn, H, W,
                                                      _stream(xz)), "ss2d_dwconv_silu_fwd")
        ctx.save_for_backward(xz, w, bb)
        ctx.Cn, ctx.wshape, ctx.has_bias = Cn, weight.shape, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        xz, w, bb = ctx.saved_tensors
        B, H, W, Cs = xz.shape
        Cn = ctx.Cn
        dout = dout.contiguous().float()
        dxz = torch.empty_like(xz) if Cs == Cn else torch.zeros_like(xz)  # dense input: every element is written
        dw = torch.zeros_like(w)
        db = torch.zeros(Cn, device=xz.device, dtype=torch.float32) if ctx.has_bias else None
        scratch = torch.empty((B, H, W, Cn), device=xz.device, dtype=torch.float32)
        with torch.cuda.device(xz.device):
            _lib.check(_lib.lib().ss2d_dwconv_silu_bwd(xz.data_ptr(), Cs, w.data_ptr(), _ptr(bb), dout.data_ptr(),
                                                      scratch.data_ptr(), dxz.data_ptr(), Cs, dw.data_ptr(), _ptr(db), B, Cn, H,
                                                      W, _stream(xz)), "ss2d_dwconv_silu_bwd")
        return dxz, dw.view(ctx.wshape), db, None


def dwconv_silu(xz, weight, bias=None, channels=None):
    return DwConvSiLUFn.apply(xz, weight, bias, channels or weight.shape[0])
