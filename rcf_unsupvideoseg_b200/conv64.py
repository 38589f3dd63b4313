"""Second layer of ``flow_feat_before_agg`` -- Conv2d(64 -> 64, 3x3, padding 1) (reference
models/flow_aggregation_head_with_residual.py:89-91) -- on the 5th-generation tensor cores (csrc/rcf_conv64.cu,
tcgen05.mma with bf16 operand splits, fp32 accumulation in tensor memory).  Channels-last fp32 in and out.  CUDA only.

Kernel-level modes of these entry points (``nprod``): 3 = three bf16 products per fp32 product (fp32-grade, ~1e-5), 2 =
activations x [W_hi | W_lo] stacked in N = 128 (two products, bf16 activations), 1 = one product; with the fp16 operand
flags (``a_f16`` / ``w_f16`` / ``f16``, nprod = 1) the 16-bit words are IEEE fp16: one product at TF32-class accuracy.
The drop-in head picks a precision LEVEL from the caller's torch settings (``default_nprod``) and maps it onto these modes
inside ``rcf_head_forward`` / ``rcf_head_backward``: level 3 -> mode 3, level 2 (torch's default, autocast) -> mode 1 with
fp16 operands and a scaled fp16 gradient, level 1 -> mode 1 with bf16 operands.
"""
from __future__ import annotations

import torch

from . import _lib

WPACK_BYTES = 9 * 16384 + 2 * (9 * 8192 + 9 * 4096)      # RCF_CONV64_WPACK_BYTES
A_F16, W_F16 = 0x100, 0x200                               # RCF_CONV64_A_F16 / RCF_CONV64_W_F16 (include/rcf_loss.h)


def default_nprod(autocast: bool = False) -> int:
    """Conv-precision level of the default head from the caller's torch settings, like the reference's own convs:
    3 (fp32-grade) when TF32 convolutions are disallowed and no autocast is active; otherwise 2 -- one product of fp16
    operands with a scaled fp16 gradient.  Under autocast the reference runs fp16 convs with a loss scaler; level 2 is the
    same arithmetic with the scale chosen per call on the device, and costs what plain bf16 (level 1) costs."""
    if autocast:
        return 2
    return 2 if torch.backends.cudnn.allow_tf32 else 3


def conv64_supported(conv: torch.nn.Conv2d) -> bool:
    return (conv.in_channels == 64 and conv.out_channels == 64 and tuple(conv.kernel_size) == (3, 3)
            and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (1, 1) and tuple(conv.dilation) == (1, 1)
            and conv.groups == 1 and conv.padding_mode == "zeros")


def pack_weights(w: torch.Tensor, transpose_flip: bool, f16: bool = False) -> torch.Tensor:
    """[64,64,3,3] fp32 -> the swizzled bf16 (hi, lo) B tiles of the kernel; transpose_flip packs the data-gradient operator.
    f16: the hi words are fp16 (use with nprod = 1 and w_f16 = True in the conv call)."""
    lib = _lib.load_library()
    assert tuple(w.shape) == (64, 64, 3, 3) and w.is_cuda
    w = w.detach().float().contiguous()
    out = torch.empty(WPACK_BYTES, dtype=torch.uint8, device=w.device)
    with _lib.device_guard(w.device):
        _lib.check(lib.rcf_conv64_pack_weights(w.data_ptr(), out.data_ptr(), int(transpose_flip) | (W_F16 if f16 else 0),
                                               _lib.raw_stream(w.device)), "rcf_conv64_pack_weights")
    return out


def pack_weights_both(w: torch.Tensor):
    """One launch: (forward image, data-gradient image) of the same [64,64,3,3] weight (views of one buffer)."""
    lib = _lib.load_library()
    assert tuple(w.shape) == (64, 64, 3, 3) and w.is_cuda
    w = w.detach().float().contiguous()
    out = torch.empty(2 * WPACK_BYTES, dtype=torch.uint8, device=w.device)
    with _lib.device_guard(w.device):
        _lib.check(lib.rcf_conv64_pack_weights(w.data_ptr(), out.data_ptr(), 2,
                                               _lib.raw_stream(w.device)), "rcf_conv64_pack_weights")
    return out[:WPACK_BYTES], out[WPACK_BYTES:]


def split_bf16(x: torch.Tensor, want_lo: bool = True):
    """fp32 tensor (any dense layout, numel % 4 == 0) -> (hi, lo) bf16 tensors of the same shape / strides, x ~ hi + lo."""
    lib = _lib.load_library()
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo = torch.empty_like(x, dtype=torch.bfloat16) if want_lo else None
    with _lib.device_guard(x.device):
        _lib.check(lib.rcf_split_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr() if want_lo else None, x.numel(),
                                      _lib.raw_stream(x.device)), "rcf_split_bf16")
    return hi, lo


def conv64_pair(x_hi: torch.Tensor, x_lo, wpack: torch.Tensor, nprod: int, a_f16: bool = False, w_f16: bool = False) -> torch.Tensor:
    """x_hi / x_lo: [N,64,H,W] bf16, channels-last memory format (x ~ hi + lo); returns fp32 channels-last.
    a_f16 / w_f16 (nprod = 1 only, both or neither: mixing fp16 with bf16 operands is an illegal tcgen05 instruction): x_hi is
    a float16 tensor / wpack was packed with f16=True."""
    lib = _lib.load_library()
    if not x_hi.is_cuda:
        raise RuntimeError("conv64: CUDA tensors required (no CPU fallback)")
    N, C, H, W = x_hi.shape
    assert C == 64 and x_hi.dtype == (torch.float16 if a_f16 else torch.bfloat16) and x_hi.is_contiguous(memory_format=torch.channels_last)
    assert nprod < 3 or (x_lo is not None and x_lo.shape == x_hi.shape and x_lo.is_contiguous(memory_format=torch.channels_last))
    out = torch.empty((N, 64, H, W), dtype=torch.float32, device=x_hi.device, memory_format=torch.channels_last)
    with _lib.device_guard(x_hi.device):
        _lib.check(lib.rcf_conv64_forward(x_hi.data_ptr(), x_lo.data_ptr() if x_lo is not None else None, wpack.data_ptr(),
                                          out.data_ptr(), N, H, W, int(nprod) | (A_F16 if a_f16 else 0) | (W_F16 if w_f16 else 0),
                                          _lib.raw_stream(x_hi.device)), "rcf_conv64_forward")
    return out


def conv64_raw(x: torch.Tensor, wpack: torch.Tensor, nprod: int) -> torch.Tensor:
    """x: [N,64,H,W] fp32 (made channels-last if it is not); returns the same shape, channels-last."""
    if not x.is_cuda:
        raise RuntimeError("conv64: CUDA tensors required (no CPU fallback)")
    if x.dtype != torch.float32 or not x.is_contiguous(memory_format=torch.channels_last):
        x = x.float().contiguous(memory_format=torch.channels_last)
    hi, lo = split_bf16(x, want_lo=(nprod == 3))
    return conv64_pair(hi, lo, wpack, nprod)


_WS_BYTES = {}


def conv64_wgrad_pair(x_hi, x_lo, g_hi, g_lo, nprod: int, f16: bool = False) -> torch.Tensor:
    """dW [64,64,3,3] fp32 from the layer input x ~ x_hi + x_lo and the output gradient g ~ g_hi + g_lo (bf16 channels-last).
    f16 (nprod = 1 only): x_hi and g_hi are float16 tensors (tcgen05 wants both operands in one format)."""
    lib = _lib.load_library()
    N, C, H, W = x_hi.shape
    assert C == 64 and g_hi.shape == x_hi.shape
    for t in (x_hi, x_lo, g_hi, g_lo):
        assert t is None or (t.dtype in (torch.bfloat16, torch.float16) and t.is_contiguous(memory_format=torch.channels_last))
    assert x_hi.dtype == g_hi.dtype == (torch.float16 if f16 else torch.bfloat16)
    dev = x_hi.device
    key = (dev.index, N, H, W)
    if key not in _WS_BYTES:
        import ctypes as C_
        nbytes = C_.c_size_t()
        _lib.check(lib.rcf_conv64_wgrad_workspace_bytes(N, H, W, C_.byref(nbytes)), "rcf_conv64_wgrad_workspace_bytes")
        _WS_BYTES[key] = nbytes.value
    ws = torch.empty(_WS_BYTES[key], dtype=torch.uint8, device=dev)
    dw = torch.empty(64, 64, 3, 3, dtype=torch.float32, device=dev)
    with _lib.device_guard(dev):
        _lib.check(lib.rcf_conv64_wgrad(x_hi.data_ptr(), x_lo.data_ptr() if x_lo is not None else None, g_hi.data_ptr(),
                                        g_lo.data_ptr() if g_lo is not None else None, dw.data_ptr(), ws.data_ptr(), N, H, W,
                                        int(nprod) | ((A_F16 | W_F16) if f16 else 0), _lib.raw_stream(dev)), "rcf_conv64_wgrad")
    return dw


def conv64_wgrad_raw(x: torch.Tensor, g: torch.Tensor, nprod: int) -> torch.Tensor:
    """fp32 convenience wrapper (splits both operands first)."""
    x = x.float().contiguous(memory_format=torch.channels_last)
    g = g.float().contiguous(memory_format=torch.channels_last)
    x_hi, x_lo = split_bf16(x, want_lo=(nprod == 3))
    g_hi, g_lo = split_bf16(g, want_lo=(nprod >= 2))
    return conv64_wgrad_pair(x_hi, x_lo, g_hi, g_lo, nprod)


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
            dw = conv64_wgrad_raw(x, g, 3 if ctx.nprod >= 2 else 1)
        return dx, dw, None


def conv64(x: torch.Tensor, w: torch.Tensor, nprod: int | None = None) -> torch.Tensor:
    """Bias-free 3x3 / padding-1 convolution 64 -> 64 channels, differentiable in x and w."""
    return _Conv64Fn.apply(x, w, default_nprod() if nprod is None else nprod)
