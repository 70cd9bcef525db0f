"""Pins the CPU oracle (oracle/ss2d_oracle.c + the torch port) against vectors produced by the
Python reference itself (oracle/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ss2d_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SCAN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "scan_*.npz")))


def _close(a, b, rtol=2e-4, atol=2e-5, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(1.0, float(np.abs(b).max()))
    err = np.abs(a - b).max()
    assert err <= atol * scale + rtol * scale, f"{what}: max err {err} (scale {scale})"


@pytest.mark.parametrize("path", SCAN_FILES, ids=[os.path.basename(p)[5:-4] for p in SCAN_FILES])
def test_c_oracle_scan_matches_reference(path):
    g = dict(np.load(path))
    meta = g["meta"]
    softplus = bool(meta[8])
    args = (g["u"], g["delta"], g["A"], g["B"], g["C"], g.get("D"), g.get("z"), g.get("delta_bias"))
    f = orc.scan_fwd(*args, delta_softplus=softplus)
    _close(f["out"], g["out"], what="out")
    _close(f["last_state"], g["last_state"], what="last_state")
    # the checkpoint tensor's last chunk carries the last state in its odd slots (test_selective_scan.py:79)
    _close(f["x"][:, :, -1, 1::2], g["last_state"], what="x[...,-1,1::2]")
    b = orc.scan_bwd(*args, g["dout"], delta_softplus=softplus)
    for k in ("du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias", "dz"):
        if k in g:
            _close(b[k], g[k], what=k)
        else:
            assert b[k] is None or k in ("dD", "ddelta_bias", "dz")


@pytest.mark.parametrize("path", SCAN_FILES, ids=[os.path.basename(p)[5:-4] for p in SCAN_FILES])
def test_torch_port_matches_reference(path):
    g = {k: torch.from_numpy(v) for k, v in np.load(path).items()}
    softplus = bool(g["meta"][8])
    ins = {k: g[k].clone().requires_grad_() for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias") if k in g}
    out, last = orc.selective_scan_ref_port(ins["u"], ins["delta"], ins["A"], ins["B"], ins["C"], ins.get("D"),
                                            ins.get("z"), ins.get("delta_bias"), softplus, True)
    _close(out.detach(), g["out"], what="out")
    _close(last.detach(), g["last_state"], what="last_state")
    out.backward(g["dout"])
    for k, name in (("u", "du"), ("delta", "ddelta"), ("A", "dA"), ("B", "dB"), ("C", "dC"), ("D", "dD"),
                    ("z", "dz"), ("delta_bias", "ddelta_bias")):
        if k in ins:
            _close(ins[k].grad, g[name], what=name)


def test_c_oracle_checkpoint_layout():
    """x:(B,Dm,ceil(L/2048),2N) holds (running prod a, h) at every 2048 boundary
    (selective_scan_fwd_kernel_oflex.cuh:163-166, selective_scan_oflex.cpp:218-220)."""
    g = dict(np.load(os.path.join(GOLDEN, "scan_twochunk.npz")))
    f = orc.scan_fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g["D"], None, g["delta_bias"], delta_softplus=True)
    assert f["x"].shape == (1, 2, 2, 8)
    head = orc.scan_fwd(g["u"][..., :2048], g["delta"][..., :2048], g["A"], g["B"][..., :2048], g["C"][..., :2048],
                        g["D"], None, g["delta_bias"], delta_softplus=True)
    np.testing.assert_allclose(f["x"][:, :, 0, 1::2], head["last_state"], rtol=1e-6, atol=1e-7)


def test_cross_scan_merge_match_reference():
    g = dict(np.load(os.path.join(GOLDEN, "cross.npz")))
    for tag in "abc":
        np.testing.assert_array_equal(orc.cross_scan(g[f"{tag}_x"]), g[f"{tag}_xs"])
        _close(orc.cross_merge(g[f"{tag}_ys"]), g[f"{tag}_y"], rtol=1e-6, atol=1e-6, what="merge")
        np.testing.assert_array_equal(orc.cross_scan_port(torch.from_numpy(g[f"{tag}_x"])).numpy(), g[f"{tag}_xs"])
        _close(orc.cross_merge_port(torch.from_numpy(g[f"{tag}_ys"])).numpy(), g[f"{tag}_y"], rtol=1e-6, atol=1e-6)


def test_dwconv_silu_matches_reference():
    g = dict(np.load(os.path.join(GOLDEN, "dwconv.npz")))
    C = g["weight"].shape[0]
    y = orc.dwconv_silu_fwd(g["xz"], g["weight"], g["bias"], C=C)
    _close(y, g["y"], rtol=1e-5, atol=1e-6, what="y")
    b = orc.dwconv_silu_bwd(g["xz"], g["weight"], g["bias"], g["dy"], C=C)
    _close(b["dx"], g["dxz"][..., :C], rtol=1e-5, atol=1e-6, what="dx")
    assert np.all(g["dxz"][..., C:] == 0)
    _close(b["dweight"], g["dweight"], rtol=1e-5, atol=1e-6, what="dweight")
    _close(b["dbias"], g["dbias"], rtol=1e-5, atol=1e-6, what="dbias")
