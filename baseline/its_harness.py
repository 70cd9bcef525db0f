"""Model-level harness (TEST / BENCH INFRASTRUCTURE, not product): drives the UNMODIFIED reference ITS model — the files
baseline/fetch_its.sh stages under baseline/_ref/its_ref/models — the way ITS/train.py and ITS/eval.py do, so that the
product (focalnet_b200.patch_ss2d) can be measured and checked underneath the reference's own model code.

What is restated here (own code, citing the reference lines it follows):
  * the training step of ITS/train.py:57-91 — multi-scale L1 + 0.1 x FFT-L1 loss, global-norm clip at 0.001, Adam;
  * the evaluation step of ITS/eval.py:33-54 — reflect padding to a multiple of 32, crop, clamp, PSNR = 10 log10(1/mse);
  * config 1 of BASELINE.json: the CPU forward with torch CrossScan / CrossMerge (vmamba_layers.py:29-71) and a
    selective_scan_ref-backed scan (test_selective_scan.py:168-234, the oracle's torch port).

Three import stubs (baseline/stubs: mamba_ssm, timm.models.layers, fvcore.nn) stand in for packages this image lacks
and the model never calls at train / eval time (DropPath and trunc_normal_ are implemented, the rest raise)."""
from __future__ import annotations

import importlib
import os
import sys
from functools import partial

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ITS_REF = os.path.join(HERE, "_ref", "its_ref")
STUBS = os.path.join(HERE, "stubs")


def available() -> bool:
    return os.path.exists(os.path.join(ITS_REF, "models", "MIMOUNet.py"))


class _DeviceRewrite(torch.overrides.TorchFunctionMode):
    """The reference constructors hard-code device='cuda' (layers.py:5, vmamba_layers.py:323,401,440,538); on a CPU-only
    host (config 1, the CPU tests) rewrite that keyword instead of editing the files."""

    def __init__(self, device):
        super().__init__()
        self.device = device

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        dev = kwargs.get("device")
        if dev is not None and str(dev).startswith("cuda"):
            kwargs["device"] = self.device
        return func(*args, **kwargs)


_REF_CUDA = None


def ref_cuda_module():
    """The reference's own oflex CUDA extension rebuilt for sm_100a (oracle/build_ref.sh), or None.  Loaded under its own
    name and cached here: `sys.modules["selective_scan_cuda_oflex"]` may hold this library's drop-in of the same name."""
    global _REF_CUDA
    if _REF_CUDA is None:
        so = os.path.join(ROOT, "oracle", "_ref", "selective_scan_cuda_oflex_ref.so")
        if not (os.path.exists(so) and torch.cuda.is_available()):
            return None
        spec = importlib.util.spec_from_file_location("selective_scan_cuda_oflex_ref", so)
        _REF_CUDA = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_REF_CUDA)
    return _REF_CUDA


_models = {}


def import_reference(variant: str = "g2"):
    """-> the reference's MIMOUNet module (g2: ITS/models/MIMOUNet.py, g4: ITS/results_1mlp_g4/code/MIMOUNet.py)."""
    if variant in _models:
        return _models[variant]
    if not available():
        raise RuntimeError("reference model files are not staged: run baseline/fetch_its.sh where /root/reference exists")
    for pkg in ("mamba_ssm", "timm", "fvcore"):
        try:
            importlib.import_module(pkg)
        except ImportError:
            if STUBS not in sys.path:
                sys.path.append(STUBS)
    if ITS_REF not in sys.path:
        sys.path.insert(0, ITS_REF)
    mod = importlib.import_module("models.MIMOUNet" if variant == "g2" else "models.MIMOUNet_g4")
    _models[variant] = mod
    return mod


def build_model(variant: str = "g2", device: str = "cuda", seed: int = 1234):
    """build_net() of the reference (MIMOUNet.py:181-182) under its own seeding (main.py:11-14)."""
    mod = import_reference(variant)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if str(device).startswith("cuda"):
        return mod.build_net()
    with _DeviceRewrite(device):
        return mod.build_net()


def ss2d_modules(model):
    return [m for m in model.modules() if hasattr(m, "forward_corev2") and hasattr(m, "forward_core")]


# ---------------------------------------------------------------------------------------------------------------
# bindings of the reference SS2D core other than the shipped v4 (Triton CrossScan + reference CUDA scan)
class _RefScanFn:
    """`SelectiveScan=` object for cross_selective_scan (vmamba_layers.py:252-253): selective_scan_ref through autograd."""

    @staticmethod
    def apply(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, nrows=1, backnrows=1, oflex=True):
        from oracle.ss2d_oracle import selective_scan_ref_port
        return selective_scan_ref_port(u, delta, A, B, C, D, None, delta_bias, delta_softplus)


def bind_cpu_reference(model) -> int:
    """Config 1: every SS2D runs torch CrossScan / CrossMerge + selective_scan_ref (no CUDA, no Triton)."""
    vml = sys.modules["models.vmamba_layers"]
    n = 0
    for m in ss2d_modules(model):
        m.forward_core = partial(m.forward_corev2, force_fp32=False, SelectiveScan=_RefScanFn, no_einsum=True,
                                 CrossScan=vml.CrossScan, CrossMerge=vml.CrossMerge)
        n += 1
    return n


def bind_reference_cuda(model, triton_cross: bool = True) -> int:
    """The reference arm on the GPU: SelectiveScanOflex on the reference's own CUDA kernels (oracle/_ref) with the shipped
    Triton CrossScan / CrossMerge (forward type v4, vmamba_layers.py:447) or their torch twins (:29-71)."""
    vml = sys.modules["models.vmamba_layers"]
    ref = ref_cuda_module()
    if ref is None:
        raise RuntimeError("oracle/_ref/selective_scan_cuda_oflex_ref.so is not built")
    vml.selective_scan_cuda_oflex = ref  # the name SelectiveScanOflex.forward / .backward resolve (vmamba_layers.py:183,193)
    n = 0
    for m in ss2d_modules(model):
        kw = dict(force_fp32=False, SelectiveScan=vml.SelectiveScanOflex, no_einsum=True)
        if triton_cross:
            kw.update(CrossScan=vml.CrossScanTriton, CrossMerge=vml.CrossMergeTriton)
        else:
            kw.update(CrossScan=vml.CrossScan, CrossMerge=vml.CrossMerge)
        m.forward_core = partial(m.forward_corev2, **kw)
        m.forward = m.forwardv2
        n += 1
    return n


# ---------------------------------------------------------------------------------------------------------------
def its_loss(pred, label):
    """ITS/train.py:64-87: L1 at three scales + 0.1 x L1 between the 2-D FFTs (real / imaginary parts stacked)."""
    label2 = F.interpolate(label, scale_factor=0.5, mode="bilinear")
    label4 = F.interpolate(label, scale_factor=0.25, mode="bilinear")
    content, fft = 0.0, 0.0
    for p, l in zip(pred, (label4, label2, label)):
        content = content + F.l1_loss(p, l)
        pf, lf = torch.fft.fft2(p, dim=(-2, -1)), torch.fft.fft2(l, dim=(-2, -1))
        fft = fft + F.l1_loss(torch.stack((pf.real, pf.imag), -1), torch.stack((lf.real, lf.imag), -1))
    return content + 0.1 * fft


def make_optimizer(model, lr: float = 1e-4):
    """ITS/train.py:16 (learning rate: main.py's default 1e-4)."""
    return torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8)


def train_step(model, optimizer, x, label, clip: float = 0.001):
    """One iteration of ITS/train.py:57-91 -> loss (device scalar)."""
    optimizer.zero_grad()
    loss = its_loss(model(x), label)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
    optimizer.step()
    return loss.detach()


def synthetic_pair(batch, h, w, device="cuda", seed=0):
    """A hazy / clear pair with the ITS statistics: clear J in [0,1] (smooth random field), hazy I = J t + A (1 - t)
    with transmission t in [0.3, 0.9] and airlight A in [0.7, 1.0] (atmospheric scattering model of the RESIDE ITS set)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    low = torch.rand(batch, 3, max(h // 16, 2), max(w // 16, 2), generator=g)
    J = F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False)
    J = (J + 0.05 * torch.randn(batch, 3, h, w, generator=g)).clamp(0, 1)
    t = 0.3 + 0.6 * torch.rand(batch, 1, 1, 1, generator=g)
    A = 0.7 + 0.3 * torch.rand(batch, 1, 1, 1, generator=g)
    I = (J * t + A * (1 - t)).clamp(0, 1)
    return I.to(device), J.to(device)


def eval_forward(model, x, factor: int = 32):
    """ITS/eval.py:33-41: reflect-pad H, W up to a multiple of `factor`, run, crop the finest output back."""
    h, w = x.shape[2], x.shape[3]
    H, W = ((h + factor) // factor) * factor, ((w + factor) // factor) * factor
    padh = H - h if h % factor != 0 else 0
    padw = W - w if w % factor != 0 else 0
    xp = F.pad(x, (0, padw, 0, padh), "reflect")
    return model(xp)[2][:, :, :h, :w]


def psnr(pred, label):
    """ITS/eval.py:47,54: 10 log10(1 / mse) of the clamped prediction, per batch (dB)."""
    return float(10 * torch.log10(1 / F.mse_loss(torch.clamp(pred, 0, 1), label)))


def ssim(x, y, data_range: float = 1.0, win_size: int = 11, sigma: float = 1.5):
    """Structural similarity with the published default recipe (Wang et al. 2004; what pytorch_msssim.ssim computes, which
    ITS/eval.py:49-52 calls and this image lacks): separable Gaussian window 11 / sigma 1.5, 'valid' filtering, K = (0.01,
    0.03), mean over channels and pixels -> one value per image."""
    coords = torch.arange(win_size, dtype=x.dtype, device=x.device) - win_size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g = (g / g.sum())
    C = x.shape[1]

    def blur(t):
        t = F.conv2d(t, g.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)
        return F.conv2d(t, g.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)

    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mx, my = blur(x), blur(y)
    sxx, syy, sxy = blur(x * x) - mx * mx, blur(y * y) - my * my, blur(x * y) - mx * my
    cs = (2 * sxy + c2) / (sxx + syy + c2)
    return (((2 * mx * my + c1) / (mx * mx + my * my + c1)) * cs).flatten(1).mean(1)


def eval_metrics(pred, label):
    """ITS/eval.py:43-54: PSNR of the clamped prediction and SSIM after average-pooling both images by
    max(1, round(min(H, W) / 256)) — on the already cropped prediction.  -> (psnr dB, mean ssim)."""
    pc = torch.clamp(pred, 0, 1)
    H, W = pc.shape[2], pc.shape[3]
    r = max(1, round(min(H, W) / 256))
    size = (int(H / r), int(W / r))
    s = ssim(F.adaptive_avg_pool2d(pc, size), F.adaptive_avg_pool2d(label, size), data_range=1.0)
    return float(10 * torch.log10(1 / F.mse_loss(pc, label))), float(s.mean())


def param_count(model) -> int:
    return sum(p.numel() for p in model.parameters())
