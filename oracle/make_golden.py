"""TEST INFRASTRUCTURE ONLY — generate golden vectors by running the *Python reference itself*.

Run in the build container (where /root/reference exists):   python oracle/make_golden.py
Writes small .npz fixtures to tests/golden/.  They are committed; /root/reference does not travel
to the GPU box, the fixtures do.

What is executed from the reference (imported, unmodified, never copied):
  * kernels/selective_scan/test_selective_scan.py : selective_scan_ref (:168-234) + torch autograd
  * ITS/models/vmamba_layers.py : CrossScan / CrossMerge (:29-71) and cross_selective_scan (:200-299)
Third-party modules the reference imports but that are absent here (mamba_ssm, timm, fvcore, the
CUDA extension modules) are replaced by empty stubs purely so that the files import; none of the stubbed
symbols is on the executed path.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference():
    for n in ("selective_scan_cuda_oflex", "selective_scan_cuda", "selective_scan_cuda_core"):
        _stub(n)
    _stub("mamba_ssm", Mamba=object)
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", DropPath=type("DropPath", (torch.nn.Identity,), {}), trunc_normal_=lambda *a, **k: None)
    _stub("fvcore")
    _stub("fvcore.nn", FlopCountAnalysis=None, flop_count_str=None, flop_count=None, parameter_count=None)

    def _load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    # test_selective_scan.py rebinds the name `selective_scan_ref` to the pip mamba_ssm CUDA kernel at import
    # time (MODE switch, :319-359), so take the pure-torch definition (:168-234) straight from its AST and
    # execute that function body, unmodified, in a namespace holding only its own imports.
    import ast
    path = f"{REF}/kernels/selective_scan/test_selective_scan.py"
    tree = ast.parse(open(path).read(), path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "selective_scan_ref")
    assert (fn.lineno, fn.end_lineno) == (168, 234), (fn.lineno, fn.end_lineno)
    import torch.nn.functional as F
    from einops import rearrange, repeat
    ns = dict(torch=torch, F=F, rearrange=rearrange, repeat=repeat)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    tss = types.SimpleNamespace(selective_scan_ref=ns["selective_scan_ref"])
    sys.path.insert(0, f"{REF}/ITS/models")
    vml = _load("ref_vmamba_layers", f"{REF}/ITS/models/vmamba_layers.py")
    return tss, vml


def scan_case(tss, name, Bn, Dm, N, L, G, has_D, has_z, has_bias, softplus, seed, squeeze_bc=False):
    if os.environ.get("ONLY") and os.environ["ONLY"] not in f"scan_{name}":
        return
    # input distributions of test_selective_scan.py:406-441
    g = torch.Generator().manual_seed(seed)
    A = (-0.5 * torch.rand(Dm, N, generator=g)).requires_grad_()
    bshape = (Bn, N, L) if squeeze_bc else (Bn, G, N, L)
    Bm = torch.randn(*bshape, generator=g).requires_grad_()
    Cm = torch.randn(*bshape, generator=g).requires_grad_()
    D = torch.randn(Dm, generator=g).requires_grad_() if has_D else None
    z = torch.randn(Bn, Dm, L, generator=g).requires_grad_() if has_z else None
    bias = (0.5 * torch.rand(Dm, generator=g)).requires_grad_() if has_bias else None
    u = torch.randn(Bn, Dm, L, generator=g).requires_grad_()
    delta = (0.5 * torch.rand(Bn, Dm, L, generator=g)).requires_grad_()
    out, last = tss.selective_scan_ref(u, delta, A, Bm, Cm, D, z=z, delta_bias=bias, delta_softplus=softplus,
                                       return_last_state=True)
    dout = torch.randn(out.shape, generator=g)
    out.backward(dout)
    rec = dict(u=u, delta=delta, A=A, B=Bm, C=Cm, dout=dout, out=out, last_state=last,
               du=u.grad, ddelta=delta.grad, dA=A.grad, dB=Bm.grad, dC=Cm.grad)
    if has_D:
        rec.update(D=D, dD=D.grad)
    if has_z:
        rec.update(z=z, dz=z.grad)
    if has_bias:
        rec.update(delta_bias=bias, ddelta_bias=bias.grad)
    rec = {k: v.detach().numpy().astype(np.float32) for k, v in rec.items()}
    rec["meta"] = np.array([Bn, Dm, N, L, G, int(has_D), int(has_z), int(has_bias), int(softplus)], np.int64)
    np.savez_compressed(os.path.join(OUT, f"scan_{name}.npz"), **rec)
    print("scan", name, {k: v.shape for k, v in rec.items() if k in ("u", "B", "out")})


def cross_cases(vml):
    if os.environ.get("ONLY") and os.environ["ONLY"] not in "cross":
        return
    g = torch.Generator().manual_seed(7)
    rec = {}
    for tag, (B, C, H, W) in dict(a=(2, 3, 5, 7), b=(1, 2, 8, 8), c=(1, 1, 1, 6)).items():
        x = torch.randn(B, C, H, W, generator=g)
        xs = vml.CrossScan.apply(x)
        ys = torch.randn(B, 4, C, H, W, generator=g)
        y = vml.CrossMerge.apply(ys)
        rec.update({f"{tag}_x": x, f"{tag}_xs": xs, f"{tag}_ys": ys, f"{tag}_y": y})
    np.savez_compressed(os.path.join(OUT, "cross.npz"), **{k: v.numpy() for k, v in rec.items()})
    print("cross", list(rec))


def fused_case(tss, vml, name, B, D, H, W, N, R, seed):
    if os.environ.get("ONLY") and os.environ["ONLY"] not in f"fused_{name}":
        return
    _fused_case(tss, vml, name, B, D, H, W, N, R, seed)


def _fused_case(tss, vml, name, B, D, H, W, N, R, seed):
    """cross_selective_scan (vmamba_layers.py:200-299) with the CPU pieces the reference itself ships:
    torch CrossScan/CrossMerge + a Function wrapping selective_scan_ref."""
    K = 4

    class RefScan:
        @staticmethod
        def apply(u, delta, A, Bm, Cm, Dv, delta_bias, delta_softplus, nrows, backnrows, ssoflex):
            return tss.selective_scan_ref(u, delta, A, Bm, Cm, Dv, None, delta_bias, delta_softplus)

    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, D, H, W, generator=g).requires_grad_()
    xw = (torch.randn(K, R + 2 * N, D, generator=g) * D ** -0.5).requires_grad_()
    dtw = ((torch.rand(K, D, R, generator=g) * 2 - 1) * R ** -0.5).requires_grad_()
    dtb = (torch.rand(K, D, generator=g) * 2 - 3).requires_grad_()
    A_logs = torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(K * D, 1)
    A_logs = (A_logs + 0.1 * torch.randn(K * D, N, generator=g)).requires_grad_()
    Ds = (1 + 0.1 * torch.randn(K * D, generator=g)).requires_grad_()
    ln = torch.nn.LayerNorm(D)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.1 * torch.randn(D, generator=g))
        ln.bias.copy_(0.1 * torch.randn(D, generator=g))
    rec = {}
    for no_einsum in (True, False):
        y = vml.cross_selective_scan(x, xw, None, dtw, dtb, A_logs, Ds, delta_softplus=True, out_norm=ln,
                                     out_norm_shape="v0", SelectiveScan=RefScan, CrossScan=vml.CrossScan,
                                     CrossMerge=vml.CrossMerge, no_einsum=no_einsum)
        rec[f"y_noeinsum{int(no_einsum)}"] = y.detach()
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    rec.update(x=x, x_proj_weight=xw, dt_projs_weight=dtw, dt_projs_bias=dtb, A_logs=A_logs, Ds=Ds,
               ln_weight=ln.weight, ln_bias=ln.bias, dy=dy, dx=x.grad, dx_proj_weight=xw.grad,
               ddt_projs_weight=dtw.grad, ddt_projs_bias=dtb.grad, dA_logs=A_logs.grad, dDs=Ds.grad,
               dln_weight=ln.weight.grad, dln_bias=ln.bias.grad)
    np.savez_compressed(os.path.join(OUT, f"fused_{name}.npz"),
                        **{k: v.detach().numpy().astype(np.float32) for k, v in rec.items()})
    print("fused", name, tuple(y.shape))


def dwconv_case():
    """SS2D.forwardv2 pre-mix (vmamba_layers.py:585-594): chunk -> permute -> depthwise conv3x3+bias -> SiLU."""
    if os.environ.get("ONLY") and os.environ["ONLY"] not in "dwconv":
        return
    g = torch.Generator().manual_seed(11)
    B, H, W, C = 2, 6, 5, 8
    xz = torch.randn(B, H, W, 2 * C, generator=g).requires_grad_()
    conv = torch.nn.Conv2d(C, C, 3, padding=1, groups=C, bias=True)
    x, _ = xz.chunk(2, dim=-1)
    y = torch.nn.functional.silu(conv(x.permute(0, 3, 1, 2).contiguous()))
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    np.savez_compressed(os.path.join(OUT, "dwconv.npz"), xz=xz.detach().numpy(), weight=conv.weight.detach().numpy(),
                        bias=conv.bias.detach().numpy(), y=y.detach().numpy(), dy=dy.numpy(),
                        dxz=xz.grad.numpy(), dweight=conv.weight.grad.numpy(), dbias=conv.bias.grad.numpy())
    print("dwconv", tuple(y.shape))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    tss, vml = load_reference()
    #                name        B  Dm  N   L    G  D      z      bias   softplus seed
    scan_case(tss, "tiny",       2, 8,  4,  37,  2, True,  False, True,  True,  0)
    scan_case(tss, "nobias",     1, 6,  3,  64,  1, False, False, False, False, 1, squeeze_bc=True)
    scan_case(tss, "zgate",      2, 4,  16, 50,  4, True,  True,  True,  True,  2)
    scan_case(tss, "n16g4",      1, 8,  16, 300, 4, True,  False, True,  True,  3)
    scan_case(tss, "twochunk",   1, 2,  4,  2100, 1, True, False, True,  True,  4)
    scan_case(tss, "n1",         2, 8,  1,  130, 2, True,  False, True,  False, 5)
    cross_cases(vml)
    fused_case(tss, vml, "small", 2, 8, 5, 7, 4, 2, 21)
    fused_case(tss, vml, "n16", 1, 12, 6, 6, 16, 3, 22)
    # L % 16 == 0 (round 2): shapes the state-lanes fused kernels (fp32, dstate 16, whole 16-step blocks) take when pinned
    fused_case(tss, vml, "n16_8x8", 1, 8, 8, 8, 16, 3, 23)
    fused_case(tss, vml, "n16_4x12", 2, 20, 4, 12, 16, 6, 24)
    dwconv_case()
