"""Times the S1 scan kernels only (fwd, bwd) with CUDA events over back-to-back launches; `warpscan` / `statelanes` on the
command line pins the kernel family (test hook ss2d_set_default_family) for an A/B comparison; `s3` adds the fused seam."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from focalnet_b200 import scan_bwd, scan_fwd
from tests._util import make_scan_inputs

def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6550.0
shapes = [(8, 192, 64, 64, "micro"), (32, 192, 128, 128, "train-L16384"), (32, 192, 32, 32, "train-L1024"), (1, 192, 120, 160, "fullres-g4")]
if len(sys.argv) > 1 and sys.argv[1] == "micro":
    shapes = shapes[:1]
from focalnet_b200 import _lib
fam = _lib.FAMILY_WARPSCAN if "warpscan" in sys.argv else (_lib.FAMILY_STATELANES if "statelanes" in sys.argv else 0)
_lib.lib().ss2d_set_default_family(fam)
print("family:", {0: "auto (by problem size)", 1: "statelanes", 2: "warpscan"}[fam])
for (B, D, H, W, tag) in shapes:
    K, N, L = 4, 16, H * W
    for dt, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        if tag != "micro" and dt != torch.float32: continue
        d = make_scan_inputs(B, K * D, N, L, K, dtype=dt)
        a = (d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])
        out, x, ckpt, _ = scan_fwd(*a, True, 1, True)
        big, bc = B * K * D * L, B * K * N * L
        tf = timeit(lambda: scan_fwd(*a, True, 1, True))
        tb = timeit(lambda: scan_bwd(*a, d["dout"], x, True, 1, ckpt=ckpt))
        bf, bb = es * (2 * big + 2 * bc) + 4 * big, es * (4 * big + 2 * bc) + 4 * big + 8 * bc
        print(f"[{tag} {str(dt)[6:]}] fwd {tf*1e6:8.1f} us ({bf/tf/1e9:7.1f} GB/s {100*bf/tf/1e9/peak:5.1f}%)  "
              f"bwd {tb*1e6:8.1f} us ({bb/tb/1e9:7.1f} GB/s {100*bb/tb/1e9/peak:5.1f}%)  "
              f"fwd+bwd {(tf+tb)*1e6:8.1f} us {100*(bf+bb)/(tf+tb)/1e9/peak:5.1f}% of {peak:.0f} GB/s")
        del d, a, out, x, ckpt
        torch.cuda.empty_cache()

if "s3" in sys.argv:  # fused seam: CrossScan -> scan -> CrossMerge in one kernel (no 4x copies)
    from focalnet_b200 import FusedCrossScanFn
    for (B, D, H, W, tag) in [(8, 192, 64, 64, "micro"), (32, 192, 128, 128, "train-L16384"), (1, 192, 120, 160, "fullres-g4")]:
        K, N, L = 4, 16, H * W
        d = make_scan_inputs(B, K * D, N, L, K)
        g = torch.Generator().manual_seed(0)
        xx = torch.randn(B, D, H, W, generator=g).cuda()
        dy = torch.randn(B, D, L, generator=g).cuda()
        fa = [t.detach().clone().requires_grad_() for t in (xx, d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"])]
        tf = timeit(lambda: FusedCrossScanFn.apply(*[t.detach() for t in fa], True))
        def fb():
            yy = FusedCrossScanFn.apply(*fa, True); yy.backward(dy)
        tfb = timeit(fb, n=10)
        big, bc, pl = B * K * D * L, B * K * N * L, B * D * L
        bf = 4 * (pl + big + 2 * bc) + 4 * pl
        bb = 4 * (pl + 2 * big + 2 * bc) + 4 * (2 * pl) + 8 * bc
        print(f"[{tag} fp32] S3 fused fwd {tf*1e6:8.1f} us ({bf/tf/1e9:7.1f} GB/s {100*bf/tf/1e9/peak:5.1f}%)  fwd+bwd (autograd, incl. zero fills) "
              f"{tfb*1e6:8.1f} us ({100*(bf+bb)/tfb/1e9/peak:5.1f}%)")
        del d, fa, xx, dy
        torch.cuda.empty_cache()
