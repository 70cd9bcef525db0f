"""Short training run of the unchanged ITS model on both kernel arms from the same initial weights and the same synthetic
stream: (i) the reference's kernels (oflex CUDA rebuilt for sm_100a + Triton CrossScan / CrossMerge, torch clip + Adam),
(ii) focalnet_b200 (patch_ss2d + FlatBucket + fused clip / Adam).  Prints the two loss curves and the PSNR of both models
on held-out synthetic pairs — the north star's "restored-image PSNR within 0.01 dB" after actual optimisation steps.
Usage: python tools/train_parity.py [steps] [batch]"""
import copy, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import its_harness as H
from focalnet_b200 import patch_ss2d
from focalnet_b200.dp import FlatBucket, FusedClipAdam, dp_train_step

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.backends.cudnn.allow_tf32 = False  # both arms: the 1x1 projections round the same way, differences are the kernels'
ref = H.build_model("g2", "cuda")
ours = copy.deepcopy(ref)
H.bind_reference_cuda(ref, triton_cross=True)
patch_ss2d(ours)
ref.train(); ours.train()
opt_ref = H.make_optimizer(ref, lr=1e-4)
bucket = FlatBucket(ours.parameters()); opt = FusedClipAdam(bucket, lr=1e-4, max_norm=0.001)
print(f"step  loss(reference kernels)  loss(focalnet_b200)  |diff|   (batch {batch}, 256x256, lr 1e-4, clip 0.001)")
for s in range(steps):
    x, J = H.synthetic_pair(batch, 256, 256, "cuda", seed=1000 + s)
    torch.manual_seed(s); torch.cuda.manual_seed_all(s)
    l_ref = float(H.train_step(ref, opt_ref, x, J))
    torch.manual_seed(s); torch.cuda.manual_seed_all(s)
    l_ours = float(dp_train_step(ours, bucket, opt, H.its_loss, x, J))
    if s % 5 == 0 or s == steps - 1:
        print(f"{s:4d}  {l_ref:.6f}                {l_ours:.6f}             {abs(l_ref - l_ours):.2e}")
ref.eval(); ours.eval()
with torch.no_grad():
    for (h, w, tag) in ((256, 256, "256x256 crops"), (460, 620, "620x460 full-res")):
        x, J = H.synthetic_pair(4, h, w, "cuda", seed=7)
        p_in = H.psnr(x, J)
        p_ref, p_ours = H.psnr(H.eval_forward(ref, x), J), H.psnr(H.eval_forward(ours, x), J)
        print(f"held-out {tag}: PSNR hazy input {p_in:.4f} dB | reference kernels {p_ref:.4f} dB | focalnet_b200 {p_ours:.4f} dB | delta {abs(p_ref - p_ours):.5f} dB")
worst = max(float((p - q).abs().max() / q.abs().max().clamp_min(1e-6)) for p, q in zip(ours.parameters(), ref.parameters()))
print(f"largest relative parameter difference after {steps} steps: {worst:.2e}")
