"""A few fused (seam S3) forward+backward passes at the microbench shape — for an ncu launch list of every kernel of the op."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focalnet_b200 import FusedCrossScanFn
from tests._util import make_scan_inputs
B, D, H, W, K, N = 8, 192, 64, 64, 4, 16
if len(sys.argv) > 1 and sys.argv[1] == "train":
    B, H, W = 32, 128, 128
L = H * W
d = make_scan_inputs(B, K * D, N, L, K)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, D, H, W, generator=g).cuda()
dy = torch.randn(B, D, L, generator=g).cuda()
fa = [t.detach().clone().requires_grad_() for t in (x, d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])]
for _ in range(3):
    torch.cuda.nvtx.range_push("fused_fwd_bwd")
    y = FusedCrossScanFn.apply(*fa, True)
    y.backward(dy)
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
