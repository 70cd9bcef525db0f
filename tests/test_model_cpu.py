"""CPU checks of the model-level harness (baseline/its_harness.py) on the UNMODIFIED reference ITS model staged by
baseline/fetch_its.sh: config 1 of BASELINE.json (forward through torch CrossScan / CrossMerge + selective_scan_ref), the
restated training step, and that patch_ss2d finds and re-binds all 12 SS2D modules without touching the model files."""
import pytest
import torch

from baseline import its_harness as H

pytestmark = pytest.mark.skipif(not H.available(), reason="reference model files not staged (baseline/fetch_its.sh)")


@pytest.fixture(scope="module")
def cpu_model():
    m = H.build_model("g2", "cpu")
    assert H.bind_cpu_reference(m) == 12
    return m


def test_reference_model_shape_and_parameter_count(cpu_model):
    assert H.param_count(cpu_model) == 2541673                      # the all-reduce payload
    mods = H.ss2d_modules(cpu_model)
    assert len(mods) == 12 and all(m.A_logs.shape == (768, 16) and m.dt_projs_weight.shape == (4, 192, 6) for m in mods)


def test_config1_cpu_forward_and_eval_padding(cpu_model):
    cpu_model.eval()
    x, J = H.synthetic_pair(1, 40, 56, "cpu", seed=1)               # not a multiple of 32: eval.py:33-37 pads by reflection
    with torch.no_grad():
        y = H.eval_forward(cpu_model, x)
    assert tuple(y.shape) == (1, 3, 40, 56) and torch.isfinite(y).all()
    assert 3.0 < H.psnr(y, J) < 60.0


def test_training_step_restatement_runs_and_updates(cpu_model):
    cpu_model.train()
    opt = H.make_optimizer(cpu_model)
    x, J = H.synthetic_pair(2, 32, 32, "cpu", seed=2)
    before = [p.detach().clone() for p in cpu_model.parameters()]
    torch.manual_seed(0)
    loss = H.train_step(cpu_model, opt, x, J)
    assert torch.isfinite(loss) and float(loss) > 0
    with_grad = [p for p in cpu_model.parameters() if p.grad is not None]
    assert len(with_grad) == len(before)                            # every parameter is on the graph
    total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in with_grad))
    assert float(total) <= 0.001 * 1.0001                           # train.py:90 clip
    assert any(not torch.equal(a, b) for a, b in zip(before, cpu_model.parameters()))
    cpu_model.eval()


def test_patch_ss2d_rebinds_every_module_of_the_unchanged_model():
    from focalnet_b200 import block_supported, cross_selective_scan, patch_ss2d, unpatch_ss2d
    m = H.build_model("g2", "cpu")
    mods = H.ss2d_modules(m)
    orig = [(mm.forward_core, mm.forward) for mm in mods]
    assert all(block_supported(mm) for mm in mods)                   # the shipped v4 configuration
    assert patch_ss2d(m) == 12
    for mm, (core, _) in zip(mods, orig):
        kw = mm.forward_core.keywords
        assert kw["cross_selective_scan"] is cross_selective_scan
        assert kw["no_einsum"] is True and kw["force_fp32"] is False  # the module's own choices survive the patch
        assert mm.forward.func.__name__ == "ss2d_forward"
    assert unpatch_ss2d(m) == 12
    assert all(mm.forward_core is core and mm.forward == fwd for mm, (core, fwd) in zip(mods, orig))


def test_eval_metrics_restatement():
    """PSNR / SSIM of ITS/eval.py:43-54: identical images -> SSIM 1 and infinite PSNR; SSIM falls with noise and is symmetric."""
    torch.manual_seed(0)
    a = torch.rand(2, 3, 64, 80)
    b = (a + 0.1 * torch.randn_like(a)).clamp(0, 1)
    s_same, s_noisy, s_sym = H.ssim(a, a), H.ssim(a, b), H.ssim(b, a)
    assert torch.allclose(s_same, torch.ones(2), atol=1e-6) and (s_noisy < 0.95).all() and torch.allclose(s_noisy, s_sym, atol=1e-6)
    p, s = H.eval_metrics(b, a)
    assert 15.0 < p < 30.0 and 0.0 < s < 1.0
