"""Tiny driver for ncu captures of the fused seam S3 (FusedCrossScanFn fwd+bwd) at the microbench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._util import make_scan_inputs
from focalnet_b200 import FusedCrossScanFn

B, D, H, W, K, N = 8, 192, 64, 64, 4, 16
L = H * W
d = make_scan_inputs(B, K * D, N, L, K)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, D, H, W, generator=g).cuda()
dy = torch.randn(B, D, L, generator=g).cuda()
fa = [t.detach().clone().requires_grad_() for t in (x, d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])]
for _ in range(2):
    y = FusedCrossScanFn.apply(*fa, True)
    y.backward(dy)
torch.cuda.synchronize()
print("ok")
