import os, sys, torch
sys.path.insert(0, "/root/repo")
from baseline import its_harness as H
from focalnet_b200 import GraphedForward, patch_ss2d
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    model = H.build_model("g4", "cuda"); patch_ss2d(model); model.eval()
    x, _ = H.synthetic_pair(1, 460, 620, "cuda", seed=9)
    with torch.no_grad():
        e1 = H.eval_forward(model, x); e2 = H.eval_forward(model, x)
        gf = GraphedForward(lambda t: H.eval_forward(model, t), x)
        g1 = gf(x).clone(); g2 = gf(x).clone()
    print(f"cudnn tf32={tf32}: eager-eager {float((e1-e2).abs().max()):.3e}  graph-graph {float((g1-g2).abs().max()):.3e}  graph-eager {float((g1-e1).abs().max()):.3e}")
