"""Dev harness: forward parity (vs C oracle, vs reference CUDA) + timing at the microbench shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._util import load_ref_cuda, make_scan_inputs, rel_err
from focalnet_b200 import scan_fwd
from oracle import ss2d_oracle as orc

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

ref = load_ref_cuda()
for (B, dim, N, L, G, dt) in [(2, 16, 16, 700, 4, torch.float32), (1, 8, 4, 2100, 2, torch.float32), (2, 24, 16, 1030, 4, torch.bfloat16),
                              (1, 8, 16, 37, 1, torch.float32), (1, 6, 3, 517, 2, torch.float16)]:
    d = make_scan_inputs(B, dim, N, L, G, dtype=dt)
    out, x, ckpt, _ = scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
    o = orc.scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], None, d["delta_bias"], True)
    print((B, dim, N, L, G, dt), "out", rel_err(out, o["out"]), "last", rel_err(x[:, :, -1, 1::2], o["last_state"]),
          "x", rel_err(x[..., 1::2], o["x"][..., 1::2]))
    if ref is not None:
        ro, rx = ref.fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
        print("   vs ref cuda: out", rel_err(out, ro), "x", rel_err(x, rx), " ref vs oracle:", rel_err(ro, o["out"]))

for dt in (torch.float32, torch.bfloat16):
    d = make_scan_inputs(8, 768, 16, 4096, 4, dtype=dt)
    f = lambda: scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
    best, med = timeit(f)
    es = 4 if dt == torch.float32 else 2
    nbytes = es * (2 * 8 * 768 * 4096 + 2 * 8 * 4 * 16 * 4096) + 4 * 8 * 768 * 4096
    print(f"ours fwd {dt}: best {best*1e3:.1f} us, median {med*1e3:.1f} us, {nbytes/best/1e6:.0f} GB/s algorithmic")
    if ref is not None:
        g = lambda: ref.fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, True)
        best, med = timeit(g)
        print(f"ref  fwd {dt}: best {best*1e3:.1f} us, median {med*1e3:.1f} us, {nbytes/best/1e6:.0f} GB/s algorithmic")
