import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import focalnet_b200._lib as L
k = sys.argv[1]
if k != "0": L.LIB_PATH = L.LIB_PATH.replace("libss2d_b200.so", f"libss2d_knock{k}.so")
from tests._util import make_scan_inputs
from focalnet_b200 import scan_fwd, scan_bwd
d = make_scan_inputs(8, 768, 16, 4096, 4)
a = (d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])
out, x, ckpt, _ = scan_fwd(*a, True, 1, True)
f = lambda: scan_bwd(*a, d["dout"], x, True, 1, ckpt=ckpt)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
print("bwd knock", k, "%.1f us" % (e0.elapsed_time(e1) / 10 * 1e3))
