"""Dev: time backward tiling variants (needs a build with SS2D_TUNE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._util import make_scan_inputs, rel_err
from focalnet_b200 import scan_fwd, scan_bwd

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

d = make_scan_inputs(8, 768, 16, 4096, 4)
args = (d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])
out, x, ckpt, _ = scan_fwd(*args, True, 1, True)
f = lambda: scan_bwd(*args, d["dout"], x, True, 1, ckpt=ckpt)
base = None
for cfg in ["", "16x12x1x6", "16x8x1x8", "16x16x1x8", "8x16x1x8"]:
    if cfg: os.environ["SS2D_BWD_CFG"] = cfg
    g = f()
    if base is None: base = g
    t = timeit(f)
    errs = max(rel_err(a, b) for a, b in zip(g[:7], base[:7]))
    print(f"cfg {cfg or 'default 8x8x2xs':16s} {t*1e3:7.1f} us   max diff vs default {errs:.2e}")
