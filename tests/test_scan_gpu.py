"""GPU parity of the selective scan (seam S1) — called through the C ABI (ctypes) like a user would.

Compared against: (1) the committed golden vectors produced by the Python reference itself,
(2) the double-precision C oracle on seeded inputs, (3) the reference's own CUDA kernels rebuilt for
sm_100a (oracle/_ref), when present.  Gates: 1e-3 relative for fp32, 1e-2 for bf16/fp16 (north star)."""
import glob
import os

import numpy as np
import pytest
import torch

from tests._util import load_ref_cuda, make_scan_inputs, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("scan_family")]
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SCAN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "scan_*.npz")))
TOL = {torch.float32: 1e-3, torch.float16: 1e-2, torch.bfloat16: 1e-2}
GRADS = ("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias", "dz")


def _run_ours(d, softplus, out_float=True):
    from focalnet_b200 import scan_bwd, scan_fwd
    out, x, ckpt, out_z = scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d.get("D"), d.get("delta_bias"), softplus, 1,
                                   out_float, z=d.get("z"))
    dout = d["dout"].to(out.dtype)
    g = scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d.get("D"), d.get("delta_bias"), dout, x, softplus, 1,
                 ckpt=ckpt, z=d.get("z"), out=out if d.get("z") is not None else None)
    res = dict(out=out_z if out_z is not None else out, last_state=x[:, :, -1, 1::2], x=x)
    res.update(dict(zip(GRADS, g)))
    return res


@pytest.mark.parametrize("path", SCAN_FILES, ids=[os.path.basename(p)[5:-4] for p in SCAN_FILES])
def test_scan_matches_reference_golden(path):
    g = {k: torch.from_numpy(v) for k, v in np.load(path).items()}
    softplus = bool(g["meta"][8])
    d = {k: g[k].cuda() for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias", "dout") if k in g}
    squeeze = d["B"].dim() == 3
    if squeeze:
        d["B"], d["C"] = d["B"].unsqueeze(1), d["C"].unsqueeze(1)
    r = _run_ours(d, softplus)
    assert rel_err(r["out"], g["out"]) < 1e-3
    assert rel_err(r["last_state"], g["last_state"]) < 1e-3
    for k in GRADS:
        if k in g:
            ours = r[k].squeeze(1) if squeeze and k in ("dB", "dC") else r[k]
            assert rel_err(ours, g[k]) < 1e-3, k


CASES = [
    # batch dim  N   L     G  dtype            D     bias  z      softplus
    (2, 16, 16, 700, 4, torch.float32, True, True, False, True),
    (1, 8, 4, 2100, 2, torch.float32, True, True, False, True),       # crosses the 2048 reference chunk
    (2, 24, 16, 1030, 4, torch.bfloat16, True, True, False, True),
    (1, 8, 16, 37, 1, torch.float32, False, False, False, False),
    (1, 6, 3, 517, 2, torch.float16, True, False, False, True),        # N not a multiple of the state block, odd L
    (2, 8, 16, 256, 2, torch.float32, True, True, True, True),         # z gate
    (1, 12, 1, 1024, 4, torch.float32, True, True, False, True),       # N = 1 (the reference test grid)
    (1, 4, 20, 300, 1, torch.float32, True, True, False, True),        # N > 16
    (3, 10, 8, 1, 2, torch.float32, True, True, False, True),          # L = 1
    (1, 8, 16, 4112, 1, torch.float32, True, True, False, True),       # aligned fast path with a short last stage
    (4, 4800, 16, 80, 4, torch.float32, True, True, False, True),      # large batch*dim: 4 states per lane forward, ragged channel tile
    (2, 9480, 16, 77, 2, torch.bfloat16, True, True, False, True),     # same, unaligned rows, 16-bit
    (2, 40, 16, 160, 4, torch.float16, True, True, True, False),       # z gate, fp16, no softplus
]


@pytest.mark.parametrize("case", CASES, ids=[f"B{c[0]}D{c[1]}N{c[2]}L{c[3]}G{c[4]}{str(c[5])[6:]}{'z' if c[8] else ''}" for c in CASES])
def test_scan_matches_oracle(case):
    from oracle import ss2d_oracle as orc
    B, dim, N, L, G, dt, hasD, hasb, hasz, sp = case
    d = make_scan_inputs(B, dim, N, L, G, dtype=dt, has_D=hasD, has_bias=hasb, has_z=hasz, seed=L)
    r = _run_ours(d, sp)
    f = orc.scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], sp)
    b = orc.scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], d["dout"], sp)
    tol = TOL[dt]
    assert rel_err(r["out"], f["out"]) < tol
    assert rel_err(r["last_state"], f["last_state"]) < tol
    assert rel_err(r["x"][..., 1::2], f["x"][..., 1::2]) < tol
    for k in GRADS:
        if b[k] is not None:
            assert rel_err(r[k], b[k]) < tol * (3 if dt != torch.float32 else 1), k


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("out_float", [True, False], ids=["oflex", "core"])
def test_scan_matches_reference_cuda(dt, out_float):
    """Same inputs through the reference's own CUDA kernels (sm_100a rebuild) and through ours."""
    ref = load_ref_cuda()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    d = make_scan_inputs(2, 96, 16, 2500, 4, dtype=dt, seed=5)
    r = _run_ours(d, True, out_float)
    ro, rx = ref.fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True, 1, out_float)
    dout = d["dout"].to(ro.dtype)
    rg = ref.bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], dout, rx, True, 1)
    tol = TOL[dt]
    assert r["out"].dtype == ro.dtype
    assert rel_err(r["out"], ro) < tol
    assert rel_err(r["x"][..., 1::2], rx[..., 1::2]) < tol
    for k, theirs in zip(GRADS[:7], rg):
        assert r[k].dtype == theirs.dtype, k
        assert rel_err(r[k], theirs) < tol * (3 if dt != torch.float32 else 1), k


def test_bwd_from_reference_x_only():
    """A caller that kept only the reference's coarse x (cloned, so the fine checkpoints are gone) still gets
    correct gradients: the library rebuilds the checkpoints (ss2d_scan_bwd_params.ckpt_scratch)."""
    from focalnet_b200 import scan_bwd, scan_fwd
    d = make_scan_inputs(1, 8, 16, 1500, 2, seed=9)
    out, x, ckpt, _ = scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], True)
    a = scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], d["dout"], x, True)
    b = scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["delta_bias"], d["dout"], x.clone(), True)
    for i in range(7):
        assert rel_err(a[i], b[i]) < 1e-5


def test_selective_scan_fn_autograd_and_strides():
    """API of record: selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)."""
    from focalnet_b200 import selective_scan_fn
    from oracle import ss2d_oracle as orc
    d = make_scan_inputs(2, 8, 16, 333, 1, seed=3, has_z=True)
    Bm, Cm = d["B"][:, 0].clone(), d["C"][:, 0].clone()          # 3-D B/C are lifted to one group
    u = d["u"].transpose(1, 2).contiguous().transpose(1, 2)       # last stride != 1 -> made contiguous inside
    leaves = [t.clone().requires_grad_() for t in (u, d["delta"], d["A"], Bm, Cm, d["D"], d["z"], d["delta_bias"])]
    out, last = selective_scan_fn(*leaves[:6], leaves[6], leaves[7], True, True)
    out.backward(d["dout"])
    f = orc.scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], True)
    b = orc.scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], d["z"], d["delta_bias"], d["dout"], True)
    assert rel_err(out, f["out"]) < 1e-3 and rel_err(last, f["last_state"]) < 1e-3
    for leaf, k in zip(leaves, ("du", "ddelta", "dA", "dB", "dC", "dD", "dz", "ddelta_bias")):
        ref = b[k][:, 0] if k in ("dB", "dC") else b[k]
        assert rel_err(leaf.grad, ref) < 1e-3, k


def test_errors_match_reference_behaviour():
    from focalnet_b200 import scan_fwd
    d = make_scan_inputs(1, 6, 4, 32, 2)
    with pytest.raises(RuntimeError):   # dim % n_groups (selective_scan_oflex.cpp:191)
        scan_fwd(d["u"][:, :5], d["delta"][:, :5], d["A"][:5], d["B"], d["C"])
    with pytest.raises(RuntimeError):   # dtype mismatch (:171)
        scan_fwd(d["u"], d["delta"].half(), d["A"], d["B"], d["C"])
    with pytest.raises(RuntimeError):   # CPU tensors: no CPU path
        scan_fwd(d["u"].cpu(), d["delta"].cpu(), d["A"].cpu(), d["B"].cpu(), d["C"].cpu())


def test_strided_operands_are_honoured():
    """Batch / channel strides are arbitrary in the reference (only the last stride must be 1,
    selective_scan_oflex.cpp:181-182,198-200): feed slices of larger buffers and a (batch, L, dim)-major u."""
    from focalnet_b200 import scan_bwd, scan_fwd
    from oracle import ss2d_oracle as orc
    B, dim, N, L, G = 2, 12, 16, 530, 2
    d = make_scan_inputs(B, dim, N, L, G, seed=17)
    big_u = torch.zeros(B, dim + 3, L + 8, device="cuda")
    big_u[:, 1:dim + 1, 4:L + 4] = d["u"]
    u = big_u[:, 1:dim + 1, 4:L + 4]                               # row starts not 16-byte aligned -> scalar path
    delta = torch.zeros(B, 2 * dim, L, device="cuda")[:, ::2]      # channel stride 2*L
    delta.copy_(d["delta"])
    Bbig = torch.zeros(B, G, N + 2, L, device="cuda")
    Bbig[:, :, 1:N + 1] = d["B"]
    Bm = Bbig[:, :, 1:N + 1]
    Cm = d["C"].transpose(0, 1).contiguous().transpose(0, 1)       # (G, B, N, L) storage
    assert not u.is_contiguous() and not delta.is_contiguous() and not Bm.is_contiguous() and not Cm.is_contiguous()
    out, x, ckpt, _ = scan_fwd(u, delta, d["A"], Bm, Cm, d["D"], d["delta_bias"], True)
    g = scan_bwd(u, delta, d["A"], Bm, Cm, d["D"], d["delta_bias"], d["dout"], x, True)
    f = orc.scan_fwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], None, d["delta_bias"], True)
    b = orc.scan_bwd(d["u"], d["delta"], d["A"], d["B"], d["C"], d["D"], None, d["delta_bias"], d["dout"], True)
    assert rel_err(out, f["out"]) < 1e-3
    for k, v in zip(GRADS[:7], g):
        assert rel_err(v, b[k]) < 1e-3, k


def test_large_model_like_shape_properties():
    """At the full microbench size the oracle is too slow to run in a unit test; check size-independent properties:
    (1) batch rows are independent (scan of a batch slice == slice of the scan), (2) linearity of out in u for the
    D=0 system, (3) sharding the channel axis by groups gives the same result, (4) determinism of the forward."""
    from focalnet_b200 import scan_fwd
    B, dim, N, L, G = 8, 768, 16, 4096, 4
    d = make_scan_inputs(B, dim, N, L, G, seed=2, model_like=True)
    a = (d["u"], d["delta"], d["A"], d["B"], d["C"])
    out = scan_fwd(*a, None, d["delta_bias"], True)[0]
    out2 = scan_fwd(*a, None, d["delta_bias"], True)[0]
    assert torch.equal(out, out2)
    sl = scan_fwd(d["u"][3:5], d["delta"][3:5], d["A"], d["B"][3:5], d["C"][3:5], None, d["delta_bias"], True)[0]
    assert torch.equal(sl, out[3:5])
    lin = scan_fwd(2.5 * d["u"], d["delta"], d["A"], d["B"], d["C"], None, d["delta_bias"], True)[0]
    assert rel_err(lin, 2.5 * out) < 1e-5
    per = dim // G
    grp = scan_fwd(d["u"][:, per:2 * per], d["delta"][:, per:2 * per], d["A"][per:2 * per], d["B"][:, 1:2], d["C"][:, 1:2],
                   None, d["delta_bias"][per:2 * per], True)[0]
    assert torch.equal(grp, out[:, per:2 * per])
    assert torch.isfinite(out).all()
