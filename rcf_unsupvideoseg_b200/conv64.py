"""Second layer of ``flow_feat_before_agg`` -- Conv2d(64 -> 64, 3x3, padding 1) (reference
models/flow_aggregation_head_with_residual.py:89-91) -- on the 5th-generation tensor cores (csrc/rcf_conv64.cu,
tcgen05.mma with bf16 operand splits, fp32 accumulation in tensor memory).  Channels-last fp32 in and out.  CUDA only.

Precision follows what the reference's own convs do under the caller's torch settings:
  * ``torch.backends.cudnn.allow_tf32 = False``  -> 3 bf16 products per fp32 product (fp32-grade, ~1e-5)
  * torch's default (TF32 convolutions allowed)  -> 2 products: weights split hi+lo, activations rounded to bf16
  * inside ``torch.autocast``                    -> 1 product (plain bf16 operands, fp32 accumulation)
"""
from __future__ import annotations

import torch

from . import _lib

WPACK_BYTES = 9 * 16384


def default_nprod(autocast: bool = False) -> int:
    if autocast:
        return 1
    return 2 if torch.backends.cudnn.allow_tf32 else 3


def conv64_supported(conv: torch.nn.Conv2d) -> bool:
    return (conv.in_channels == 64 and conv.out_channels == 64 and tuple(conv.kernel_size) == (3, 3)
            and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (1, 1) and tuple(conv.dilation) == (1, 1)
            and conv.groups == 1 and conv.padding_mode == "zeros")


def pack_weights(w: torch.Tensor, transpose_flip: bool) -> torch.Tensor:
    """[64,64,3,3] fp32 -> the swizzled bf16 (hi, lo) B tiles of the kernel; transpose_flip packs the data-gradient operator."""
    lib = _lib.load_library()
    assert tuple(w.shape) == (64, 64, 3, 3) and w.is_cuda
    w = w.detach().float().contiguous()
    out = torch.empty(WPACK_BYTES, dtype=torch.uint8, device=w.device)
    with _lib.device_guard(w.device):
        _lib.check(lib.rcf_conv64_pack_weights(w.data_ptr(), out.data_ptr(), int(transpose_flip),
                                               torch.cuda.current_stream(w.device).cuda_stream), "rcf_conv64_pack_weights")
    return out


def conv64_raw(x: torch.Tensor, wpack: torch.Tensor, nprod: int) -> torch.Tensor:
    """x: [N,64,H,W] fp32 in channels-last memory format; returns the same shape / format."""
    lib = _lib.load_library()
    if not x.is_cuda:
        raise RuntimeError("conv64: CUDA tensors required (no CPU fallback)")
    N, C, H, W = x.shape
    assert C == 64
    if x.dtype != torch.float32 or not x.is_contiguous(memory_format=torch.channels_last):
        x = x.float().contiguous(memory_format=torch.channels_last)
    out = torch.empty((N, 64, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    with _lib.device_guard(x.device):
        _lib.check(lib.rcf_conv64_forward(x.data_ptr(), wpack.data_ptr(), out.data_ptr(), N, H, W, int(nprod),
                                          torch.cuda.current_stream(x.device).cuda_stream), "rcf_conv64_forward")
    return out


class _Conv64Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, nprod):
        ctx.nprod = nprod
        ctx.save_for_backward(x, w)
        return conv64_raw(x, pack_weights(w, False), nprod)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = conv64_raw(g, pack_weights(w, True), ctx.nprod)
        if ctx.needs_input_grad[1]:
            g_cl = g.contiguous(memory_format=torch.channels_last)
            dw = torch.ops.aten.convolution_backward(g_cl, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1,
                                                     (False, True, False))[1]
        return dx, dw, None


def conv64(x: torch.Tensor, w: torch.Tensor, nprod: int | None = None) -> torch.Tensor:
    """Bias-free 3x3 / padding-1 convolution 64 -> 64 channels, differentiable in x and w."""
    return _Conv64Fn.apply(x, w, default_nprod() if nprod is None else nprod)
