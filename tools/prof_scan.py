"""Tiny driver for ncu captures: runs the microbench-shaped scan fwd (and bwd with --bwd) a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._util import make_scan_inputs
from focalnet_b200 import scan_fwd, scan_bwd

dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
d = make_scan_inputs(8, 768, 16, 4096, 4, dtype=dt)
for _ in range(3):
    out, x, ckpt, _ = scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
    if "--bwd" in sys.argv:
        scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], d["dout"], x, True, 1, ckpt=ckpt)
torch.cuda.synchronize()
print("ok")
