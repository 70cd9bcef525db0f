"""ncu driver: one launch of each small (HBM-bound) kernel at the training shape (B=32, 192 channels, 128x128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focalnet_b200 import cross_merge, cross_scan, dwconv_silu, merge_norm_gate
B, D, H, W = (32, 192, 128, 128) if "micro" not in sys.argv else (8, 192, 64, 64)
L = H * W
x = torch.randn(B, D, H, W, device="cuda")
xs = cross_scan(x); y = cross_merge(xs, H, W)
xh = torch.randn(B, H, W, D, device="cuda", requires_grad=True)
w = torch.randn(D, 1, 3, 3, device="cuda", requires_grad=True); b = torch.randn(D, device="cuda", requires_grad=True)
o = dwconv_silu(xh, w, b, D); o.backward(torch.randn_like(o))
ym = torch.randn(B, D, L, device="cuda", requires_grad=True)
zz = torch.randn(B, H, W, D, device="cuda", requires_grad=True)
lw, lb = torch.ones(D, device="cuda", requires_grad=True), torch.zeros(D, device="cuda", requires_grad=True)
om = merge_norm_gate(ym, lw, lb, 1e-5, z=zz); om.backward(torch.randn_like(om))
from focalnet_b200.ss2d import DtProjFn
from focalnet_b200 import _lib
xd = torch.randn(B, 4, 38, L, device="cuda", requires_grad=True)
Wd = torch.randn(4, D, 6, device="cuda", requires_grad=True)
od = DtProjFn.apply(xd[:, :, :6], Wd); od.backward(torch.randn_like(od))
yt = torch.empty(B * D, W, H, device="cuda")
_lib.lib().ss2d_plane_transpose(x.data_ptr(), yt.data_ptr(), B * D, H, W, 0, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print("ok")
