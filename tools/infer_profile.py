"""torch.profiler kernel table of the g4 full-resolution inference leg (8 images of 620x460 per GPU) with patch_ss2d."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import its_harness as H
from focalnet_b200 import patch_ss2d
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = H.build_model("g4", "cuda"); model.eval(); patch_ss2d(model)
x, J = H.synthetic_pair(T, 460, 620, "cuda", seed=1)
with torch.no_grad():
    for _ in range(3):
        H.eval_forward(model, x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        H.eval_forward(model, x)
    t_issue = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 5
    print(f"host issue time {t_issue*1e3:.2f} ms / batch, wall {t_all*1e3:.2f} ms / batch (T={T})")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            H.eval_forward(model, x)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
