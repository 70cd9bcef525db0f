"""Drop-in for the reference's pybind module ``selective_scan_cuda_oflex``
(kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_oflex.cpp:157-363; imported by name at
ITS/models/vmamba_layers.py:74-79 and called at :183,:193).  Put this directory on PYTHONPATH ahead of (or instead
of) the reference's built extension: same ``fwd`` / ``bwd`` signatures, shapes, dtypes and RuntimeErrors; the work
is done by libss2d_b200.so through the C ABI.  No CPU path."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from focalnet_b200.selective_scan import scan_bwd as _bwd, scan_fwd as _fwd  # noqa: E402


def fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, out_float=True):
    """-> [out, x]; out fp32 if out_float else u.dtype; x (batch, dim, ceil(L/2048), 2*dstate) fp32."""
    out, x, _ckpt, _ = _fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, out_float)
    return [out, x]


def bwd(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, nrows):
    """-> [du, ddelta, dA, dB, dC, dD, ddelta_bias]"""
    return list(_bwd(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, nrows)[:7])
