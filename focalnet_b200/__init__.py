"""focalnet_b200 — B200-native (sm_100a) SS2D hot path of the FocalNet/VMamba ITS dehazing model.

Only the path of SURVEY.md §8 lives here: hand-written CUDA kernels + their C ABI (``csrc/``,
``include/ss2d_b200.h``) and the host-side mirror of the reference's operator interfaces.
"""
from . import _lib  # noqa: F401
from .graphs import GraphedForward  # noqa: F401
from .selective_scan import build_selective_scan_fn, scan_bwd, scan_fwd, selective_scan_fn  # noqa: F401
from .ss2d import (CrossMerge, CrossMergeTriton, CrossScan, CrossScanTriton, FusedCrossScanFn, SelectiveScanCore,  # noqa: F401
                   SelectiveScanOflex, block_supported, cross_merge, cross_scan, cross_selective_scan, dwconv_silu,
                   merge_norm_gate, patch_ss2d, ss2d_forward, unpatch_ss2d)

__all__ = ["selective_scan_fn", "build_selective_scan_fn", "scan_fwd", "scan_bwd", "cross_scan", "cross_merge", "CrossScan",
           "CrossMerge", "CrossScanTriton", "CrossMergeTriton", "SelectiveScanOflex", "SelectiveScanCore",
           "cross_selective_scan", "FusedCrossScanFn", "dwconv_silu", "merge_norm_gate", "ss2d_forward", "patch_ss2d",
           "unpatch_ss2d", "block_supported", "GraphedForward"]
