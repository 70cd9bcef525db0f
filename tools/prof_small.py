"""ncu driver: one launch of each small (HBM-bound) kernel at the microbench and training shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focalnet_b200 import cross_merge, cross_scan, dwconv_silu
for (B, D, H, W) in [(8, 192, 64, 64), (32, 192, 128, 128)]:
    x = torch.randn(B, D, H, W, device="cuda")
    xs = cross_scan(x); y = cross_merge(xs, H, W)
    xz = torch.randn(B, H, W, 2 * D, device="cuda", requires_grad=True)
    w = torch.randn(D, 1, 3, 3, device="cuda", requires_grad=True); b = torch.randn(D, device="cuda", requires_grad=True)
    o = dwconv_silu(xz, w, b, D); o.backward(torch.randn_like(o))
torch.cuda.synchronize(); print("ok")
