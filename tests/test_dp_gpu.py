"""GPU parity of the optimizer side of the training step: ss2d_optim_clip_adam over a FlatBucket against
torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (what ITS/train.py:89-91 calls), and one full dp_train_step of the
unchanged model against the harness's restatement of the reference step."""
import copy

import pytest
import torch

from tests._util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("max_norm", [0.001, 10.0, 0.0], ids=["clip", "noclip-active", "clip-off"])
def test_fused_clip_adam_matches_torch(max_norm):
    from focalnet_b200.dp import FlatBucket, FusedClipAdam
    torch.manual_seed(0)
    shapes = [(33, 7), (5,), (64, 3, 3, 3), (1,), (129,)]
    ours = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    theirs = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    b = FlatBucket(ours, segments=2)
    opt = FusedClipAdam(b, lr=1e-3, max_norm=max_norm)
    ref = torch.optim.Adam(theirs, lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for step in range(4):
        grads = [torch.randn_like(p) * (0.1 if step % 2 else 3.0) for p in ours]
        for p, q, g in zip(ours, theirs, grads):
            b.grad_view(p).copy_(g)
            q.grad = g.clone()
        norm_ref = torch.nn.utils.clip_grad_norm_(theirs, max_norm) if max_norm > 0 else None
        ref.step()
        opt.step()
        if norm_ref is not None:
            assert rel_err(opt.last_norm, norm_ref.view(1)) < 1e-5
        for p, q in zip(ours, theirs):
            assert rel_err(p, q) < 1e-5
            assert float(b.grad_view(p).abs().max()) == 0.0   # zero_grad folded into the pass
    assert rel_err(opt.exp_avg[: ours[-1].numel()], ref.state[theirs[-1]]["exp_avg"].flatten()) < 1e-5  # reversed order


def test_dp_train_step_matches_reference_step_on_the_unchanged_model():
    from baseline import its_harness as H
    if not H.available():
        pytest.skip("reference model files not staged")
    from focalnet_b200 import patch_ss2d
    from focalnet_b200.dp import FlatBucket, FusedClipAdam, dp_train_step
    model = H.build_model("g2", "cuda")
    assert patch_ss2d(model) == 12
    twin = copy.deepcopy(model)
    patch_ss2d(twin)          # deepcopy keeps the bound partials of `model`: re-bind to the copy's own modules
    model.train(); twin.train()
    x, J = H.synthetic_pair(4, 128, 128, "cuda", seed=8)
    opt_ref = H.make_optimizer(twin, lr=1e-4)
    bucket = FlatBucket(model.parameters())
    opt = FusedClipAdam(bucket, lr=1e-4, max_norm=0.001)
    for step in range(2):
        torch.manual_seed(step); torch.cuda.manual_seed_all(step)
        opt_ref.zero_grad()
        l_ref = H.its_loss(twin(x), J)
        l_ref.backward()
        torch.manual_seed(step); torch.cuda.manual_seed_all(step)
        bucket.begin_step()
        l_ours = H.its_loss(model(x), J)
        l_ours.backward()
        bucket.finish_reduce()
        assert abs(float(l_ref) - float(l_ours)) <= 1e-5 * abs(float(l_ref))
        gmax = max(float(q.grad.abs().max()) for q in twin.parameters())
        for (n, p), q in zip(model.named_parameters(), twin.parameters()):
            assert rel_err(p.grad, q.grad, floor=1e-2 * gmax) < 1e-3, n  # = rtol 1e-3 + atol 1e-5 * gmax: analytically-zero gradients hold rounding noise
            bucket.grad_view(p).copy_(q.grad)   # identical inputs for the optimizer comparison (Adam's m / sqrt(v) amplifies noise-level grads)
        torch.nn.utils.clip_grad_norm_(twin.parameters(), 0.001)
        opt_ref.step()
        opt.step()
        for (n, p), q in zip(model.named_parameters(), twin.parameters()):
            assert rel_err(p, q, floor=1e-6) < 1e-5, n
    # and the one-call form
    torch.manual_seed(9)
    assert torch.isfinite(dp_train_step(model, bucket, opt, H.its_loss, x, J))


def test_optimizer_state_round_trips_with_torch_adam_checkpoints(tmp_path):
    """ITS/train.py:110-113 saves {'model', 'optimizer': torch.optim.Adam.state_dict(), 'epoch'} and :24-27 restores it: the
    fused optimizer reads and writes the same dictionary, so a run can move between the reference's optimizer and this one."""
    from focalnet_b200.dp import FlatBucket, FusedClipAdam
    torch.manual_seed(3)
    shapes = [(17, 5), (9,), (4, 3, 3, 3)]
    theirs = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    ref = torch.optim.Adam(theirs, lr=2e-3, betas=(0.9, 0.999), eps=1e-8)
    for _ in range(3):
        for q in theirs:
            q.grad = torch.randn_like(q)
        ref.step()
    path = str(tmp_path / "model.pkl")
    torch.save({"model": [q.detach().clone() for q in theirs], "optimizer": ref.state_dict(), "epoch": 3}, path)
    ck = torch.load(path, weights_only=False)
    ours = [torch.nn.Parameter(t.clone()) for t in ck["model"]]
    b = FlatBucket(ours)
    opt = FusedClipAdam(b, lr=1.0, max_norm=0.0)
    opt.load_state_dict(ck["optimizer"])
    assert opt.steps == 3 and opt.lr == 2e-3
    for _ in range(2):  # continue on both sides from identical gradients
        gs = [torch.randn_like(q) for q in theirs]
        for p, q, g in zip(ours, theirs, gs):
            b.grad_view(p).copy_(g)
            q.grad = g.clone()
        ref.step()
        opt.step()
    for p, q in zip(ours, theirs):
        assert rel_err(p, q) < 1e-5
    back = torch.optim.Adam([torch.nn.Parameter(t.detach().clone()) for t in ours], lr=1.0)
    back.load_state_dict(opt.state_dict())           # and back into the reference's optimizer
    assert back.param_groups[0]["lr"] == 2e-3 and int(back.state[back.param_groups[0]["params"][0]]["step"]) == 5
    assert rel_err(back.state[back.param_groups[0]["params"][1]]["exp_avg"], ref.state[theirs[1]]["exp_avg"]) < 1e-5
