"""Bilinear resize (input staging of the motion loss) as one sm_100a launch per call, forward and backward.

Reference call sites: `F.interpolate(all_pred_residual, self.mask_size, mode='bilinear')`
(models/flow_aggregation_head_with_residual.py:271-273, :294-296) and mmseg's `resize(gt_*_flows, size=..., mode='bilinear',
align_corners=...)` in the caller (models/rcf_model.py:438-442).  Same arithmetic as ATen's upsample_bilinear2d; the
backward is a deterministic gather (csrc/rcf_resize.cu).  Several tensors of identical shape (the forward and backward
residual maps) go through ONE launch.  CUDA only: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import torch

from . import _lib


def _ptrs(ts):
    return (C.c_void_p * 2)(*[t.data_ptr() for t in ts], *([None] * (2 - len(ts))))


class _ResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, size: Tuple[int, int], align_corners: bool, *xs):
        lib = _lib.load_library()
        assert 1 <= len(xs) <= 2
        x0 = xs[0]
        if not x0.is_cuda:
            raise RuntimeError("resize_bilinear: CUDA tensors required (no CPU fallback)")
        assert x0.dim() == 4 and all(x.shape == x0.shape and x.device == x0.device for x in xs)
        B, Cc, h, w = x0.shape
        H, W = int(size[0]), int(size[1])
        xv = [x.detach().float().contiguous() for x in xs]
        outs = [torch.empty(B, Cc, H, W, dtype=torch.float32, device=x0.device) for _ in xs]
        with _lib.device_guard(x0.device):
            _lib.check(lib.rcf_resize_bilinear_forward(_ptrs(xv), _ptrs(outs), len(xs), B * Cc, h, w, H, W,
                                                       int(bool(align_corners)),
                                                       _lib.raw_stream(x0.device)),
                       "rcf_resize_bilinear_forward")
        ctx.meta = (B, Cc, h, w, H, W, bool(align_corners), len(xs))
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        lib = _lib.load_library()
        B, Cc, h, w, H, W, align, n = ctx.meta
        idx = [i for i, g in enumerate(gs) if g is not None and ctx.needs_input_grad[2 + i]]
        res = [None] * n
        if idx:
            gv = [gs[i].float().contiguous() for i in idx]
            gi = [torch.empty(B, Cc, h, w, dtype=torch.float32, device=gv[0].device) for _ in idx]
            with _lib.device_guard(gv[0].device):
                _lib.check(lib.rcf_resize_bilinear_backward(_ptrs(gv), _ptrs(gi), len(idx), B * Cc, h, w, H, W, int(align),
                                                            _lib.raw_stream(gv[0].device)),
                           "rcf_resize_bilinear_backward")
            for i, g in zip(idx, gi):
                res[i] = g
        return (None, None, *res)


def resize_bilinear_multi(xs: Sequence[torch.Tensor], size, align_corners: bool = False):
    """Resize 1 or 2 equally shaped [B,C,h,w] tensors to `size` = (H, W) in one launch; returns a tuple."""
    return _ResizeFn.apply((int(size[0]), int(size[1])), bool(align_corners), *xs)


def resize_bilinear(x: torch.Tensor, size, align_corners: bool = False) -> torch.Tensor:
    """Drop-in for F.interpolate(x, size, mode='bilinear', align_corners=align_corners) on fp32 CUDA tensors."""
    out = resize_bilinear_multi([x], size, align_corners)[0]
    return out if out.dtype == x.dtype else out.to(x.dtype)      # F.interpolate preserves the input dtype


def stage_flow_hwc(flow_hwc: torch.Tensor, size, align_corners: bool = False, channel_scale=None) -> torch.Tensor:
    """RAFT flow as the loader holds it, [N, h, w, 2] (HWC; dataset/data.py:114-133) -> [N, 2, H, W] at `size`:
    optional per-channel scale (FlowTransform.scale_flow), the HWC -> CHW transpose (dataset/transforms.py:850) and the
    bilinear resize of models/rcf_model.py:438-442 in one kernel.  No gradient (ground truth)."""
    lib = _lib.load_library()
    if not flow_hwc.is_cuda:
        raise RuntimeError("stage_flow_hwc: CUDA tensors required (no CPU fallback)")
    assert flow_hwc.dim() == 4 and flow_hwc.shape[-1] <= 4, "flow [N, h, w, C<=4]"
    x = flow_hwc.detach().float().contiguous()
    N, h, w, Cc = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty(N, Cc, H, W, dtype=torch.float32, device=x.device)
    cs = (C.c_float * Cc)(*[float(v) for v in channel_scale]) if channel_scale is not None else None
    with _lib.device_guard(x.device):
        _lib.check(lib.rcf_flow_stage_hwc(x.data_ptr(), out.data_ptr(), N, Cc, h, w, H, W, int(bool(align_corners)), cs,
                                          _lib.raw_stream(x.device)), "rcf_flow_stage_hwc")
    return out
