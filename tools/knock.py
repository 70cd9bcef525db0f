import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import focalnet_b200._lib as L
k = sys.argv[1]
if k != "0": L.LIB_PATH = L.LIB_PATH.replace("libss2d_b200.so", f"libss2d_knock{k}.so")
from tests._util import make_scan_inputs, rel_err
from focalnet_b200 import scan_fwd
from oracle import ss2d_oracle as orc
ds = make_scan_inputs(2, 16, 16, 700, 4, seed=1)
o = scan_fwd(ds["u"], ds["delta"], ds["A"], ds["B"], ds["C"], ds["D"], ds["delta_bias"], True, 1, True)[0]
f = orc.scan_fwd(ds["u"], ds["delta"], ds["A"], ds["B"], ds["C"], ds["D"], None, ds["delta_bias"], True)
d = make_scan_inputs(8, 768, 16, 4096, 4)
fn = lambda: scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
for _ in range(5): fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
print("variant", k, "fwd %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3), "err vs oracle %.2e" % rel_err(o, f["out"]))
