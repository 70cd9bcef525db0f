"""Dev harness: fwd+bwd timing at the microbench shape, ours vs the reference CUDA rebuild (back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests._util import load_ref_cuda, make_scan_inputs, rel_err
from focalnet_b200 import scan_fwd, scan_bwd

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

ref = load_ref_cuda()
for dt in (torch.float32, torch.bfloat16):
    d = make_scan_inputs(8, 768, 16, 4096, 4, dtype=dt)
    args = (d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])
    out, x, ckpt, _ = scan_fwd(*args, True, 1, True)
    tf = timeit(lambda: scan_fwd(*args, True, 1, True))
    tb = timeit(lambda: scan_bwd(*args, d["dout"], x, True, 1, ckpt=ckpt))
    print(f"ours {dt}: fwd {tf*1e3:.1f} us  bwd {tb*1e3:.1f} us  total {1e3*(tf+tb):.1f} us")
    if ref is not None:
        ro, rx = ref.fwd(*args, True, 1, True)
        rf = timeit(lambda: ref.fwd(*args, True, 1, True))
        rb = timeit(lambda: ref.bwd(*args, d["dout"], rx, True, 1))
        print(f"ref  {dt}: fwd {rf*1e3:.1f} us  bwd {rb*1e3:.1f} us  total {1e3*(rf+rb):.1f} us")
