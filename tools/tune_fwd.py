"""Dev: time forward tiling variants (needs a build with SS2D_TUNE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._util import make_scan_inputs, rel_err
from focalnet_b200 import scan_fwd

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts)//2]

d = make_scan_inputs(8, 768, 16, 4096, 4)
f = lambda: scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
base = None
for cfg in ["", "16x4x1x4s4", "16x8x1x2s4", "16x4x1x2s8", "8x4x1x4s8", "8x4x2x4s8", "8x4x1x6s4", "8x2x1x8s4"]:
    if cfg: os.environ["SS2D_FWD_CFG"] = cfg
    out = f()[0]
    if base is None: base = out
    best, med = timeit(f)
    print(f"cfg {cfg or 'default 16x8x2':16s} best {best*1e3:7.1f} us  median {med*1e3:7.1f} us   diff vs default {rel_err(out, base):.2e}")
