"""Times every kernel of the library at the microbench shape (B=8, D=192, K=4, N=16, 64x64) with CUDA events over
back-to-back launches, and prints achieved algorithmic GB/s.  Output is committed under profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focalnet_b200 import (CrossMerge, CrossScan, FusedCrossScanFn, cross_merge, cross_scan, dwconv_silu, scan_bwd, scan_fwd)
from tests._util import make_scan_inputs

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
rows = []
def rep(name, secs, nbytes):
    rows.append((name, secs * 1e6, nbytes / 1e6, nbytes / secs / 1e9, nbytes / secs / 1e9 / peak))
    print(f"{name:44s} {secs*1e6:9.1f} us  {nbytes/1e6:8.1f} MB  {nbytes/secs/1e9:8.1f} GB/s  {100*nbytes/secs/1e9/peak:5.1f}% of {peak:.0f}")

for (B, D, H, W, tag) in [(8, 192, 64, 64, "micro"), (32, 192, 128, 128, "train-L16384"), (1, 192, 120, 160, "fullres-g4")]:
    K, N, L = 4, 16, H * W
    for dt, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        if tag != "micro" and dt != torch.float32: continue
        g = torch.Generator().manual_seed(0)
        d = make_scan_inputs(B, K * D, N, L, K, dtype=dt)
        a = (d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])
        out, x, ckpt, _ = scan_fwd(*a, True, 1, True)
        big, bc = B * K * D * L, B * K * N * L
        nm = f"[{tag} {str(dt)[6:]}]"
        rep(f"{nm} S1 scan fwd", timeit(lambda: scan_fwd(*a, True, 1, True)), es * (2 * big + 2 * bc) + 4 * big)
        rep(f"{nm} S1 scan bwd", timeit(lambda: scan_bwd(*a, d["dout"], x, True, 1, ckpt=ckpt)), es * (4 * big + 2 * bc) + 4 * big + 8 * bc)
        xx = torch.randn(B, D, H, W, generator=g).to(dt).cuda()
        dy = torch.randn(B, D, L, generator=g).cuda()
        fa = [t.detach().clone().requires_grad_() for t in (xx, d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])]
        y = FusedCrossScanFn.apply(*fa, True)
        rep(f"{nm} S3 fused fwd (no 4x copies)", timeit(lambda: FusedCrossScanFn.apply(*[t.detach() for t in fa], True)),
            es * (B * D * L + big + 2 * bc) + 4 * B * D * L)
        def fb():
            yy = FusedCrossScanFn.apply(*fa, True); yy.backward(dy)
        t_fb = timeit(fb, n=10)
        rep(f"{nm} S3 fused fwd+bwd (autograd)", t_fb, es * (B * D * L + big + 2 * bc) + 4 * B * D * L + es * (B * D * L + 2 * big + 2 * bc) + 4 * (2 * B * D * L) + 8 * bc)
        if dt == torch.float32 or tag == "micro":
            xs = cross_scan(xx)
            rep(f"{nm} S2 cross_scan", timeit(lambda: cross_scan(xx)), 5 * B * D * L * es)
            rep(f"{nm} S2 cross_merge", timeit(lambda: cross_merge(xs, H, W)), 5 * B * D * L * es)
    xz = torch.randn(B, H, W, 2 * D).cuda().requires_grad_()
    w = torch.randn(D, 1, 3, 3).cuda().requires_grad_(); bb = torch.randn(D).cuda().requires_grad_()
    rep(f"[{tag} float32] dwconv3x3+bias+SiLU fwd", timeit(lambda: dwconv_silu(xz.detach(), w.detach(), bb.detach(), D)), 8 * B * D * L)
    yy = dwconv_silu(xz, w, bb, D); gg = torch.randn_like(yy)
    def db():
        y2 = dwconv_silu(xz, w, bb, D); y2.backward(gg)
    rep(f"[{tag} float32] dwconv fwd+bwd (autograd)", timeit(db, n=10), 8 * B * D * L + 4 * B * D * L * 5)
    # library ops the reference uses for the same pre-mix (permute copy + cuDNN depthwise conv + SiLU)
    conv = torch.nn.Conv2d(D, D, 3, padding=1, groups=D).cuda()
    def refmix():
        return torch.nn.functional.silu(conv(xz.detach()[..., :D].permute(0, 3, 1, 2).contiguous()))
    rep(f"[{tag} float32] reference pre-mix (ATen+cuDNN)", timeit(refmix), 8 * B * D * L)
    # dense d_inner-channel input (what ss2d_forward feeds since round 2: in_proj split into x-half / z-half GEMMs)
    xh = torch.randn(B, H, W, D).cuda().requires_grad_()
    rep(f"[{tag} float32] dwconv fwd, dense input", timeit(lambda: dwconv_silu(xh.detach(), w.detach(), bb.detach(), D)), 8 * B * D * L)
    def db2():
        y2 = dwconv_silu(xh, w, bb, D); torch.autograd.grad(y2, (xh, w, bb), gg)
    rep(f"[{tag} float32] dwconv fwd+bwd, dense input", timeit(db2, n=10), 8 * B * D * L + 4 * B * D * L * 5)
    from focalnet_b200 import merge_norm_gate
    ym = torch.randn(B, D, L).cuda().requires_grad_()
    zz = torch.randn(B, H, W, D).cuda().requires_grad_()
    lw, lb = torch.ones(D).cuda().requires_grad_(), torch.zeros(D).cuda().requires_grad_()
    rep(f"[{tag} float32] merge+LayerNorm+gate fwd", timeit(lambda: merge_norm_gate(ym.detach(), lw.detach(), lb.detach(), 1e-5, z=zz.detach())), 12 * B * D * L)
    go = torch.randn(B, L, D).cuda()
    def mb():
        o = merge_norm_gate(ym, lw, lb, 1e-5, z=zz); torch.autograd.grad(o, (ym, lw, lb, zz), go)
    rep(f"[{tag} float32] merge+LayerNorm+gate fwd+bwd", timeit(mb, n=10), 12 * B * D * L + 20 * B * D * L)
    # the two projections of the fused core (vmamba_layers.py:262-264): library calls vs this library's dt_proj stream kernels
    import torch.nn.functional as Fn
    from focalnet_b200.ss2d import DtProjFn
    R = 6
    xc = torch.randn(B, D, L).cuda()
    Wx = torch.randn(K * (R + 2 * N), D, 1).cuda() * D ** -0.5
    rep(f"[{tag} float32] x_proj as conv1d (cuDNN, TF32 default)", timeit(lambda: Fn.conv1d(xc, Wx)), 4 * (B * D * L + B * K * (R + 2 * N) * L))
    rep(f"[{tag} float32] x_proj as matmul (cuBLAS fp32)", timeit(lambda: torch.matmul(Wx[:, :, 0], xc)), 4 * (B * D * L + B * K * (R + 2 * N) * L))
    xd = torch.randn(B, K, R + 2 * N, L).cuda().requires_grad_()
    Wd = torch.randn(K, D, R).cuda().requires_grad_()
    gd = torch.randn(B, K * D, L).cuda()
    rep(f"[{tag} float32] dt_proj fwd (ss2d_dt_proj_fwd)", timeit(lambda: DtProjFn.apply(xd.detach()[:, :, :R], Wd.detach())), 4 * B * K * L * (R + D))
    rep(f"[{tag} float32] dt_proj fwd, library grouped conv1d", timeit(lambda: Fn.conv1d(xd.detach()[:, :, :R].reshape(B, K * R, L), Wd.detach().reshape(K * D, R, 1), groups=K)), 4 * B * K * L * (R + D))
    def dtb(ours):
        o = DtProjFn.apply(xd[:, :, :R], Wd) if ours else Fn.conv1d(xd[:, :, :R].reshape(B, K * R, L), Wd.reshape(K * D, R, 1), groups=K)
        torch.autograd.grad(o, (xd, Wd), gd)
    rep(f"[{tag} float32] dt_proj fwd+bwd (ours, autograd)", timeit(lambda: dtb(True), n=10), 4 * B * K * L * (3 * D + 3 * R))
    rep(f"[{tag} float32] dt_proj fwd+bwd (library conv1d, autograd)", timeit(lambda: dtb(False), n=10), 4 * B * K * L * (3 * D + 3 * R))
    del xc, xd, gd
    torch.cuda.empty_cache()
