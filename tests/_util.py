"""Shared helpers for the parity tests (test infrastructure)."""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "selective_scan_cuda_oflex_ref.so")


def load_ref_cuda():
    """The reference's own oflex CUDA extension rebuilt for sm_100a (oracle/build_ref.sh), or None."""
    if not os.path.exists(REF_SO):
        return None
    spec = importlib.util.spec_from_file_location("selective_scan_cuda_oflex_ref", REF_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_scan_inputs(batch, dim, N, L, G, dtype=torch.float32, device="cuda", seed=0, has_D=True, has_bias=True,
                     has_z=False, model_like=False):
    """Input distributions of the reference test (test_selective_scan.py:406-441); `model_like` switches to the
    SS2D initialisation range (A_n = -n, softplus(delta+bias) in [1e-3, 0.1]; vmamba_layers.py:510-552)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if model_like:
        A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(dim, 1) * (1 + 0.05 * torch.rand(dim, N, generator=g))
    else:
        A = -0.5 * torch.rand(dim, N, generator=g)
    Bm = torch.randn(batch, G, N, L, generator=g)
    Cm = torch.randn(batch, G, N, L, generator=g)
    D = torch.randn(dim, generator=g) if has_D else None
    z = torch.randn(batch, dim, L, generator=g) if has_z else None
    if model_like:
        dt = torch.exp(torch.rand(dim, generator=g) * (np.log(0.1) - np.log(1e-3)) + np.log(1e-3))
        bias = (dt + torch.log(-torch.expm1(-dt))) if has_bias else None
        delta = 0.5 * torch.randn(batch, dim, L, generator=g)
    else:
        bias = 0.5 * torch.rand(dim, generator=g) if has_bias else None
        delta = 0.5 * torch.rand(batch, dim, L, generator=g)
    u = torch.randn(batch, dim, L, generator=g)
    dout = torch.randn(batch, dim, L, generator=g)
    cast = lambda t: None if t is None else t.to(device=device, dtype=dtype)
    f32 = lambda t: None if t is None else t.to(device=device, dtype=torch.float32)
    return dict(u=cast(u), delta=cast(delta), A=f32(A), B=cast(Bm), C=cast(Cm), D=f32(D), z=cast(z),
                delta_bias=f32(bias), dout=f32(dout))


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(max |b|, floor) — the 'relative' of the north star's 1e-3 / 1e-2 gates.  The floor keeps
    an identically-zero reference (e.g. dA when L == 1) from turning fp32 rounding noise into an infinite ratio."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(floor))
