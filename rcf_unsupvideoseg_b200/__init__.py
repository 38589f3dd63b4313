"""B200-native (sm_100a) RCF relaxed-common-fate motion loss.

Public surface (mirrors /root/reference/models/flow_aggregation_head_with_residual.py):
    FlowAggregationHeadWithResidual, get_norm_flow, Objectview
plus the functional layer over the C ABI (include/rcf_loss.h):
    LossSpec, rcf_motion_loss, RcfMotionLossFn, load_library
"""
from ._lib import LIB_PATH, RcfLibraryError, load_library  # noqa: F401
from .function import LossSpec, RcfMotionLossFn, rcf_motion_loss  # noqa: F401
from .head import FlowAggregationHeadWithResidual, Objectview, get_norm_flow  # noqa: F401
from . import loss_utils, warp_utils  # noqa: F401  (mirrors of the reference's utils/loss_utils.py, utils/warp_utils.py)
from . import mask_ops, resize  # noqa: F401  (caller-side staging: fused softmax+entropy, bilinear resize)

__all__ = ["FlowAggregationHeadWithResidual", "get_norm_flow", "Objectview", "LossSpec", "rcf_motion_loss",
           "RcfMotionLossFn", "load_library", "RcfLibraryError", "LIB_PATH"]
