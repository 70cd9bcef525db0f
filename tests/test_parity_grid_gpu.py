"""GPU parity at the sizes that are actually benched and shipped (round-1 verdict, "close the parity holes"):

* the microbench configuration of bench.py (B=8, dim 768, N=16, L=4096, G=4; fp32 and bf16), DEFAULT kernel dispatch — no
  family pin — against BOTH the double-precision C oracle and the reference's own CUDA kernels (oracle/_ref): out, the
  reference checkpoint tensor x (even slots = running prod a, odd slots = h) and all seven gradients;
* SURVEY §4's shape grid (dim 768, G 4, L in {1024, 1200, 4800, 16384, 19200}) against the reference CUDA and the oracle;
* the fused SS2D core (CrossScan / CrossMerge in the addressing), forward AND backward, against the C-oracle composition
  cross_scan -> scan_fwd / scan_bwd -> cross_merge, with each kernel family pinned in turn.

Gates (north star): 1e-3 relative fp32, 1e-2 bf16 (gradients of 16-bit runs: 3e-2, the reference test's own slack for
reduced precision, test_selective_scan.py:398-401)."""
import pytest
import torch

from tests._util import load_ref_cuda, make_scan_inputs, rel_err

pytestmark = pytest.mark.gpu
GRADS = ("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias")
TOL = {torch.float32: 1e-3, torch.bfloat16: 1e-2}


def _ours(d, out_float=True):
    from focalnet_b200 import scan_bwd, scan_fwd
    out, x, ckpt, _ = scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, out_float)
    g = scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], d["dout"].to(out.dtype), x, True, 1, ckpt=ckpt)
    torch.cuda.synchronize()
    return out, x, ckpt, dict(zip(GRADS, g))


def _check_vs_ref_cuda(d, out, x, grads, dt, out_float=True):
    ref = load_ref_cuda()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    ro, rx = ref.fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, out_float)
    rg = ref.bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], d["dout"].to(ro.dtype), rx, True, 1)
    tol = TOL[dt]
    assert out.dtype == ro.dtype and tuple(x.shape) == tuple(rx.shape)
    assert rel_err(out, ro) < tol
    assert rel_err(x[..., 1::2], rx[..., 1::2]) < tol, "x odd slots (h at the end of every 2048-step chunk)"
    assert rel_err(x[..., 0::2], rx[..., 0::2]) < tol, "x even slots (running prod a)"
    for k, theirs in zip(GRADS, rg):
        assert grads[k].dtype == theirs.dtype, k
        assert rel_err(grads[k], theirs) < tol * (3 if dt != torch.float32 else 1), k


def _check_vs_oracle(d, out, x, grads, dt):
    from oracle import ss2d_oracle as orc
    f = orc.scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], None, d["delta_bias"], True)
    b = orc.scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], None, d["delta_bias"], d["dout"], True)
    tol = TOL[dt]
    assert rel_err(out, f["out"]) < tol
    assert rel_err(x[..., 1::2], f["x"][..., 1::2]) < tol and rel_err(x[..., 0::2], f["x"][..., 0::2]) < tol
    for k in GRADS:
        assert rel_err(grads[k], b[k]) < tol * (3 if dt != torch.float32 else 1), k


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("model_like", [False, True], ids=["testdist", "modeldist"])
def test_benched_shape_matches_oracle_and_reference_cuda(dt, model_like):
    """bench.py's configuration, default dispatch (state-lanes, 2 states per lane, FAST, whole 16-channel tiles, 384 CTAs)."""
    from focalnet_b200 import _lib
    d = make_scan_inputs(8, 768, 16, 4096, 4, dtype=dt, seed=0, model_like=model_like)
    out, x, ckpt, grads = _ours(d)
    assert ckpt.family == _lib.FAMILY_STATELANES  # what the bench line times
    _check_vs_oracle(d, out, x, grads, dt)
    _check_vs_ref_cuda(d, out, x, grads, dt)


GRID = [  # the shapes the g2 training model and the g4 full-resolution model put through the scan
    (32, 768, 16, 1024, torch.float32),
    (2, 768, 16, 16384, torch.float32),
    (1, 768, 16, 19200, torch.float32),
    (1, 768, 16, 4800, torch.float32),
    (1, 768, 16, 1200, torch.float32),
    (8, 768, 16, 4800, torch.bfloat16),
    (1, 768, 16, 19200, torch.bfloat16),
]


@pytest.mark.parametrize("case", GRID, ids=[f"B{c[0]}L{c[3]}{str(c[4])[6:]}" for c in GRID])
def test_model_shape_grid_matches_oracle_and_reference_cuda(case):
    B, dim, N, L, dt = case
    d = make_scan_inputs(B, dim, N, L, 4, dtype=dt, seed=L, model_like=(L % 3 == 0))
    out, x, ckpt, grads = _ours(d)
    _check_vs_oracle(d, out, x, grads, dt)
    _check_vs_ref_cuda(d, out, x, grads, dt)


@pytest.mark.parametrize("family", ["statelanes", "warpscan"])
@pytest.mark.parametrize("shape", [(2, 192, 64, 64), (1, 192, 120, 160), (1, 24, 30, 40)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("deterministic", [False, True], ids=["", "det"])
def test_fused_core_fwd_bwd_matches_oracle_composition(shape, family, deterministic):
    """FusedCrossScanFn forward and backward == cross_scan -> selective scan -> cross_merge of the C oracle."""
    from focalnet_b200 import FusedCrossScanFn, _lib
    from oracle import ss2d_oracle as orc
    B, D, H, W = shape
    N, L = 16, H * W
    g = torch.Generator().manual_seed(L + D)
    x = torch.randn(B, D, H, W, generator=g).cuda()
    delta = (0.5 * torch.rand(B, 4 * D, L, generator=g)).cuda()
    A = (-0.5 * torch.rand(4 * D, N, generator=g)).cuda()
    Bs, Cs = torch.randn(B, 4, N, L, generator=g).cuda(), torch.randn(B, 4, N, L, generator=g).cuda()
    Ds, bias = torch.randn(4 * D, generator=g).cuda(), (0.5 * torch.rand(4 * D, generator=g)).cuda()
    dy = torch.randn(B, D, L, generator=g).cuda()
    fam = {"statelanes": _lib.FAMILY_STATELANES, "warpscan": _lib.FAMILY_WARPSCAN}[family]
    old = _lib.lib().ss2d_set_default_family(fam)
    try:
        expect = fam if (L % 16 == 0 or fam == _lib.FAMILY_WARPSCAN) else _lib.FAMILY_WARPSCAN
        assert _lib.lib().ss2d_cross_family(B, D, H, W, N, _lib.F32, 0) == expect
        leaves = [t.clone().requires_grad_() for t in (x, delta, A, Bs, Cs, Ds, bias)]
        y = FusedCrossScanFn.apply(*leaves, True, deterministic)
        y.backward(dy)
        if deterministic:  # bit-reproducible y and dx (csm_triton.py:45-80 is deterministic too)
            leaves2 = [t.clone().requires_grad_() for t in (x, delta, A, Bs, Cs, Ds, bias)]
            y2 = FusedCrossScanFn.apply(*leaves2, True, True)
            y2.backward(dy)
            assert torch.equal(y, y2) and torch.equal(leaves[0].grad, leaves2[0].grad)
    finally:
        _lib.lib().ss2d_set_default_family(old)
    xs = torch.from_numpy(orc.cross_scan(x)).reshape(B, 4 * D, L)
    f = orc.scan_fwd(xs, delta, A, Bs, Cs, Ds, None, bias, True)
    y_ref = orc.cross_merge(f["out"].reshape(B, 4, D, H, W))
    assert rel_err(y, y_ref) < 1e-3
    # backward of CrossMerge = CrossScan of dy; backward of CrossScan = CrossMerge of du (csm_triton.py:177-185,202-210)
    dys = torch.from_numpy(orc.cross_scan(dy.view(B, D, H, W))).reshape(B, 4 * D, L)
    b = orc.scan_bwd(xs, delta, A, Bs, Cs, Ds, None, bias, dys, True)
    dx_ref = orc.cross_merge(torch.as_tensor(b["du"]).reshape(B, 4, D, H, W))
    assert rel_err(leaves[0].grad.reshape(B, D, L), dx_ref) < 1e-3, "dx"
    for leaf, k in zip(leaves[1:], ("ddelta", "dA", "dB", "dC", "dD", "ddelta_bias")):
        assert rel_err(leaf.grad, b[k]) < 1e-3, k


@pytest.mark.parametrize("family", ["statelanes", "warpscan"])
def test_fused_core_recompute_mode_and_inference_without_checkpoints(family):
    """`FusedCrossScanFn.recompute` (states rebuilt in the backward instead of kept) gives the gradients of the default
    mode; a no-grad forward (no checkpoint buffer at all) gives the same y."""
    from focalnet_b200 import FusedCrossScanFn, _lib
    B, D, H, W, N = 2, 48, 32, 48, 16
    L = H * W
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, D, H, W, generator=g).cuda()
    delta = (0.5 * torch.rand(B, 4 * D, L, generator=g)).cuda()
    A = (-0.5 * torch.rand(4 * D, N, generator=g)).cuda()
    Bs, Cs = torch.randn(B, 4, N, L, generator=g).cuda(), torch.randn(B, 4, N, L, generator=g).cuda()
    Ds, bias = torch.randn(4 * D, generator=g).cuda(), (0.5 * torch.rand(4 * D, generator=g)).cuda()
    dy = torch.randn(B, D, L, generator=g).cuda()
    fam = {"statelanes": _lib.FAMILY_STATELANES, "warpscan": _lib.FAMILY_WARPSCAN}[family]
    old = _lib.lib().ss2d_set_default_family(fam)
    try:
        res = []
        for rec in (False, True):
            FusedCrossScanFn.recompute = rec
            leaves = [t.clone().requires_grad_() for t in (x, delta, A, Bs, Cs, Ds, bias)]
            y = FusedCrossScanFn.apply(*leaves, True)
            y.backward(dy)
            res.append((y.detach(), [t.grad for t in leaves]))
        with torch.no_grad():
            y_inf = FusedCrossScanFn.apply(x, delta, A, Bs, Cs, Ds, bias, True)
    finally:
        FusedCrossScanFn.recompute = False
        _lib.lib().ss2d_set_default_family(old)
    assert rel_err(res[1][0], res[0][0]) < 1e-6 and rel_err(y_inf, res[0][0]) < 1e-6
    for a, b, k in zip(res[1][1], res[0][1], ("dx", "ddelta", "dA", "dB", "dC", "dD", "dbias")):
        assert rel_err(a, b) < 1e-5, k
