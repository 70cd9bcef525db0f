"""Tiny driver for ncu captures: runs the scan fwd (and bwd with --bwd) a few times.
   python tools/prof_scan.py [--bwd] [--bf16] [--shape B,L]   (default: the microbench shape 8,4096)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._util import make_scan_inputs
from focalnet_b200 import scan_fwd, scan_bwd

dt = torch.bfloat16 if "--bf16" in sys.argv else torch.float32
B, L = 8, 4096
if "--shape" in sys.argv:
    B, L = (int(v) for v in sys.argv[sys.argv.index("--shape") + 1].split(","))
d = make_scan_inputs(B, 768, 16, L, 4, dtype=dt)
for _ in range(3):
    out, x, ckpt, _ = scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
    if "--bwd" in sys.argv:
        scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], d["dout"], x, True, 1, ckpt=ckpt)
torch.cuda.synchronize()
print("ok")
