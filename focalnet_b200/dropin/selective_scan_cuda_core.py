"""Drop-in for the reference's pybind module ``selective_scan_cuda_core`` (cus/selective_scan.cpp; imported at
ITS/models/vmamba_layers.py:81-86, called at :161,:171): identical to the oflex module except that ``out`` always
has the input dtype (no ``out_float`` argument)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from focalnet_b200.selective_scan import scan_bwd as _bwd, scan_fwd as _fwd  # noqa: E402


def fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows):
    out, x, _ckpt, _ = _fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows, False)
    return [out, x]


def bwd(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, nrows):
    return list(_bwd(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, nrows)[:7])
