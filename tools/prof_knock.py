import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import focalnet_b200._lib as L
k = sys.argv[1]
if k != "0": L.LIB_PATH = L.LIB_PATH.replace("libss2d_b200.so", f"libss2d_knock{k}.so")
from tests._util import make_scan_inputs
from focalnet_b200 import scan_fwd
d = make_scan_inputs(8, 768, 16, 4096, 4)
for _ in range(3): scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
torch.cuda.synchronize(); print("ok")
