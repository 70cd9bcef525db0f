"""Kernel-level breakdown (torch.profiler) of one SS2D block forward+backward on this library's path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
sys.argv = [sys.argv[0]]
import importlib.util
spec = importlib.util.spec_from_file_location("bb", os.path.join(os.path.dirname(__file__), "bench_block.py"))
# reuse the module definition without running the benchmark loop
src = open(os.path.join(os.path.dirname(__file__), "bench_block.py")).read().split("torch.manual_seed(0)")[0]
ns = {"__name__": "bb", "__file__": __file__}
exec(compile(src, "bench_block_defs", "exec"), ns)
from focalnet_b200 import ss2d_forward
torch.manual_seed(0)
m = ns["SS2DShaped"]().cuda()
B, H, W = (8, 64, 64) if len(sys.argv) < 2 else (32, 128, 128)
x = torch.randn(B, H, W, ns["dm"], device="cuda", requires_grad=True)
g = torch.randn(B, H, W, ns["dm"], device="cuda")
def step(f):
    m.zero_grad(set_to_none=True)
    f(m, x).backward(g)
for which, f in (("ours", ss2d_forward), ("reference-structured", ns["reference_structured"])):
    for _ in range(3): step(f)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): step(f)
        torch.cuda.synchronize()
    print(f"==== {which}: B={B} {H}x{W}, 5 iterations fwd+bwd")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
