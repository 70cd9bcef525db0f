"""Where does the training step spend its device time?  torch.profiler kernel table of one data-parallel step of the
unchanged ITS model with patch_ss2d (ours) or on the reference's kernels (ref).  Usage: python tools/train_profile.py [ours|ref] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import its_harness as H  # noqa: E402

arm = sys.argv[1] if len(sys.argv) > 1 else "ours"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
model = H.build_model("g2", "cuda")
if arm == "ours":
    from focalnet_b200 import patch_ss2d
    from focalnet_b200.dp import FlatBucket, FusedClipAdam, dp_train_step
    patch_ss2d(model)
    bucket = FlatBucket(model.parameters())
    opt = FusedClipAdam(bucket)
    step = lambda: dp_train_step(model, bucket, opt, H.its_loss, x, J)
else:
    H.bind_reference_cuda(model, triton_cross=True)
    opt = H.make_optimizer(model)
    step = lambda: H.train_step(model, opt, x, J)
model.train()
x, J = H.synthetic_pair(batch, 256, 256, "cuda", seed=1)
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
