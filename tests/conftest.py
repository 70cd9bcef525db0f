import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(params=["statelanes", "warpscan"])
def scan_family(request):
    """Runs a GPU test once per scan-kernel family.  The library picks the family by problem size (state-lanes from
    ~4.6 k channel sequences of dstate 16, warp-scan otherwise); the test hook ss2d_set_default_family pins what "auto"
    means so that the small parity shapes exercise both (shapes a family does not cover, e.g. dstate != 16, fall through
    to the other one)."""
    from focalnet_b200 import _lib
    fam = {"statelanes": _lib.FAMILY_STATELANES, "warpscan": _lib.FAMILY_WARPSCAN}[request.param]
    old = _lib.lib().ss2d_set_default_family(fam)
    yield request.param
    _lib.lib().ss2d_set_default_family(old)


@pytest.fixture(autouse=True)
def _strict_fp32_library_ops():
    """The fused SS2D core keeps x_proj / dt_proj on the reference's library call (1x1 conv1d through cuDNN), which
    follows torch.backends.cudnn.allow_tf32 (default True, as in the reference's own runs).  Parity tests compare with
    fp64 compositions at 1e-3, so they pin the library ops to true fp32."""
    import torch
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old
