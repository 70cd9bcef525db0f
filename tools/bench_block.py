"""Whole SS2D block (SS2D.forwardv2 of the ITS model: d_model 96, d_inner 192, N 16, R 6, K 4), forward + backward:
this library's fused path (ss2d_forward) vs the reference-structured composition (permute + cuDNN depthwise conv +
SiLU, materialised CrossScan, grouped projections, the reference's own scan kernels rebuilt for sm_100a, CrossMerge,
transpose + LayerNorm, gate) — CUDA events over back-to-back iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from focalnet_b200 import CrossMerge, CrossScan, ss2d_forward
from tests._util import load_ref_cuda

ref = load_ref_cuda()
dm, D, N, R, K = 96, 192, 16, 6, 4


class SS2DShaped(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.in_proj = torch.nn.Linear(dm, 2 * D, bias=False)
        self.conv2d = torch.nn.Conv2d(D, D, 3, padding=1, groups=D)
        self.x_proj_weight = torch.nn.Parameter(torch.randn(K, R + 2 * N, D) * D ** -0.5)
        self.dt_projs_weight = torch.nn.Parameter((torch.rand(K, D, R) * 2 - 1) * R ** -0.5)
        self.dt_projs_bias = torch.nn.Parameter(torch.rand(K, D) * 2 - 4)
        self.A_logs = torch.nn.Parameter(torch.log(torch.arange(1, N + 1.0)).repeat(K * D, 1))
        self.Ds = torch.nn.Parameter(torch.ones(K * D))
        self.out_norm = torch.nn.LayerNorm(D)
        self.out_proj = torch.nn.Linear(D, dm, bias=False)
        self.dropout = torch.nn.Identity()


class RefScan(torch.autograd.Function):  # SelectiveScanOflex of vmamba_layers.py:177-196 on the reference's kernels
    @staticmethod
    def forward(ctx, u, delta, A, B, C, Dv, bias):
        out, x = ref.fwd(u, delta, A, B, C, Dv, bias, True, 1, True)
        ctx.save_for_backward(u, delta, A, B, C, Dv, bias, x)
        return out

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, B, C, Dv, bias, x = ctx.saved_tensors
        return tuple(ref.bwd(u, delta, A, B, C, Dv, bias, dout.contiguous(), x, True, 1))


def reference_structured(m, x, cross_scan=CrossScan, cross_merge=CrossMerge):
    xz = m.in_proj(x)
    xx, z = xz.chunk(2, dim=-1)
    z = F.silu(z)
    xx = F.silu(m.conv2d(xx.permute(0, 3, 1, 2).contiguous()))
    B, _, H, W = xx.shape
    L = H * W
    xs = cross_scan.apply(xx)
    x_dbl = F.conv1d(xs.view(B, -1, L), m.x_proj_weight.view(-1, D, 1), groups=K)
    dts, Bs, Cs = torch.split(x_dbl.view(B, K, -1, L), [R, N, N], dim=2)
    dts = F.conv1d(dts.contiguous().view(B, -1, L), m.dt_projs_weight.view(K * D, -1, 1), groups=K)
    ys = RefScan.apply(xs.view(B, -1, L), dts.contiguous(), -torch.exp(m.A_logs.float()), Bs.contiguous(), Cs.contiguous(),
                       m.Ds.float(), m.dt_projs_bias.view(-1).float())
    y = cross_merge.apply(ys.view(B, K, -1, H, W))
    y = m.out_norm(y.transpose(1, 2).contiguous()).view(B, H, W, -1)
    return m.out_proj(y * z)


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


torch.manual_seed(0)
m = SS2DShaped().cuda()
for (B, H, W) in [(8, 64, 64), (32, 128, 128), (32, 64, 64), (1, 120, 160)]:
    x = torch.randn(B, H, W, dm, device="cuda", requires_grad=True)
    g = torch.randn(B, H, W, dm, device="cuda")

    def run(f):
        def step():
            m.zero_grad(set_to_none=True)
            out = f(m, x)
            out.backward(g)
        return step
    with torch.no_grad():
        f_ours = timeit(lambda: ss2d_forward(m, x))
    t_ours = timeit(run(ss2d_forward))
    line = f"SS2D block B={B} {H}x{W}: ours fwd {f_ours:7.3f} ms  fwd+bwd {t_ours:7.3f} ms"
    if ref is not None:
        with torch.no_grad():
            f_ref = timeit(lambda: reference_structured(m, x))
        t_ref = timeit(run(reference_structured))
        with torch.no_grad():
            err = float((ss2d_forward(m, x) - reference_structured(m, x)).abs().max() / reference_structured(m, x).abs().max())
        line += f" | reference-structured (ref scan kernels, sm_100a rebuild) fwd {f_ref:7.3f} ms  fwd+bwd {t_ref:7.3f} ms"
        line += f" | speed-up fwd {f_ref / f_ours:4.2f}x  fwd+bwd {t_ref / t_ours:4.2f}x | max rel diff {err:.1e}"
    print(line, flush=True)
