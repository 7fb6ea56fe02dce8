"""nndepth_b200 -- the stereo-correlation hot path of anhtu293/nndepth as hand-written sm_100a kernels.

Public surface = the reference's own class names; everything runs through
the C ABI in ``include/nndepth_b200.h`` (``libnndepth_b200.so``).  No CPU fallback.
"""
from ._lib import NNDepthError, build_library, load as load_library  # noqa: F401
from .corr import (CorrBlock1D, GroupCorrBlock1D, linear_sampler, lookup_indices,  # noqa: F401
                   set_volume_precision, get_volume_precision)
from .igev import GeometryAwareCostVolume, soft_argmin  # noqa: F401
from .agcl import AGCL  # noqa: F401
from .upsample import convex_upsample  # noqa: F401

__all__ = [
    "CorrBlock1D", "GroupCorrBlock1D", "linear_sampler", "lookup_indices", "set_volume_precision",
    "get_volume_precision", "GeometryAwareCostVolume", "soft_argmin", "AGCL", "convex_upsample", "NNDepthError", "build_library",
    "load_library",
]
