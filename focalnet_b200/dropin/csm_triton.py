"""Drop-in for ITS/models/csm_triton.py (the reference's Triton CrossScan / CrossMerge): same class names and
``.apply`` signatures, implemented by the CUDA kernels of libss2d_b200.so — no Triton.  vmamba_layers.py:23-26
falls back to ``from csm_triton import ...`` when the relative import fails, so placing this directory on
PYTHONPATH and removing/renaming the package-local csm_triton.py is enough."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from focalnet_b200.ss2d import CrossMerge as CrossMergeTriton, CrossScan as CrossScanTriton  # noqa: E402,F401


class CrossScanTriton1b1:  # imported by vmamba_layers.py:24 but never used by the ITS model
    @staticmethod
    def apply(*a, **k):
        raise NotImplementedError("CrossScanTriton1b1 is unused by the ITS model and not provided")
