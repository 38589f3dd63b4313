"""First layer of ``flow_feat_before_agg`` -- Conv2d(2 -> Cf, ks, pad (ks-1)//2) + LeakyReLU (reference :84-88) -- as one
hand-written kernel each way (csrc/rcf_stem.cu): clamp + conv + bias + LeakyReLU forward, LeakyReLU-backward + weight
gradient + bias gradient backward.  Output is channels-last, which the second (cuDNN tensor-core) conv consumes natively.
No gradient flows to the RAFT flow.  CUDA only."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib


def stem_supported(Cf: int, ks: int) -> bool:
    return ks in (1, 3, 5) and Cf >= 4 and Cf % 4 == 0 and Cf <= 128 and 256 % (Cf // 4) == 0


def _views(flows: Sequence[torch.Tensor]):
    out = []
    for f in flows:
        f = f.detach()
        if f.dtype != torch.float32:
            f = f.float()
        s = f.stride()
        if not (s[3] == 1 and s[2] == f.shape[3] and s[1] == f.shape[2] * f.shape[3]):
            f = f.contiguous()
        out.append(f)
    return out


def stem_forward_raw(fl, w, b, clamp_t: float, slope: float):
    """fl: prepared per-direction flows (see _views); returns (act channels-last [ndir*B,Cf,H,W], sign bits or None)."""
    lib = _lib.load_library()
    ndir = len(fl)
    B, two, H, W = fl[0].shape
    assert two == 2 and all(f.shape == fl[0].shape for f in fl)
    if not fl[0].is_cuda:
        raise RuntimeError("stem: CUDA tensors required (no CPU fallback)")
    Cf, _, ks, _ = w.shape
    dev = fl[0].device
    act = torch.empty((ndir * B, Cf, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
    # Cf == 64: tensor-core kernels; the forward leaves one sign bit per (pixel, channel) for the backward (8 B/px)
    sign = torch.empty(ndir * B * H * W * 2, dtype=torch.int32, device=dev) if (Cf == 64 and 0.0 <= slope <= 1.0) else None
    ptrs = (C.c_void_p * 2)(*[f.data_ptr() for f in fl], *([None] * (2 - ndir)))
    strides = (C.c_int64 * 2)(*[f.stride(0) for f in fl], *([0] * (2 - ndir)))
    with _lib.device_guard(dev):
        _lib.check(lib.rcf_stem_forward(ptrs, strides, ndir, B, H, W, Cf, ks, w.data_ptr(), b.data_ptr(),
                                        float(clamp_t), float(slope), act.data_ptr(),
                                        sign.data_ptr() if sign is not None else None,
                                        _lib.raw_stream(dev)), "rcf_stem_forward")
    return act, sign


def stem_backward_raw(fl, w_shape, clamp_t: float, slope: float, act, sign, dact, nprod: int = 3):
    """(dw, db) of the stem from dact (gradient w.r.t. its output) and either the sign bits or the activation map."""
    lib = _lib.load_library()
    ndir = len(fl)
    B, _, H, W = fl[0].shape
    Cf, _, ks, _ = w_shape
    dev = dact.device
    g = dact.float().contiguous(memory_format=torch.channels_last)
    nbytes = C.c_size_t()
    _lib.check(lib.rcf_stem_workspace_bytes(ndir, B, H, W, Cf, ks, C.byref(nbytes)), "rcf_stem_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    dw = torch.empty(tuple(w_shape), dtype=torch.float32, device=dev)
    db = torch.empty(Cf, dtype=torch.float32, device=dev)
    ptrs = (C.c_void_p * 2)(*[f.data_ptr() for f in fl], *([None] * (2 - ndir)))
    strides = (C.c_int64 * 2)(*[f.stride(0) for f in fl], *([0] * (2 - ndir)))
    with _lib.device_guard(dev):
        _lib.check(lib.rcf_stem_backward(ptrs, strides, ndir, B, H, W, Cf, ks, clamp_t, slope,
                                         None if sign is not None else act.data_ptr(),
                                         sign.data_ptr() if sign is not None else None, g.data_ptr(), dw.data_ptr(),
                                         db.data_ptr(), ws.data_ptr(), int(nprod), _lib.raw_stream(dev)),
                   "rcf_stem_backward")
    return dw, db


class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, bias, clamp_t: float, slope: float, *flows):
        fl = _views(flows)
        w = weight.detach().float().contiguous()
        b = bias.detach().float().contiguous()
        act, sign = stem_forward_raw(fl, w, b, clamp_t, slope)
        ctx.has_sign = sign is not None
        ctx.save_for_backward(sign if sign is not None else act, *fl)
        ctx.meta = (tuple(w.shape), float(clamp_t), float(slope), len(fl))
        return act

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dact):
        w_shape, clamp_t, slope, ndir = ctx.meta
        kept, *fl = ctx.saved_tensors          # sign bits (Cf == 64) or the activation map
        dw, db = stem_backward_raw(fl, w_shape, clamp_t, slope, None if ctx.has_sign else kept,
                                   kept if ctx.has_sign else None, dact)
        return (dw, db, None, None, *([None] * ndir))


def flow_stem(flows: Sequence[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor, clamp_t: Optional[float],
              slope: float = 0.1) -> torch.Tensor:
    """LeakyReLU(conv(clamp(flow)) + bias) for 1 or 2 directions of [B,2,H,W] flows -> channels-last [ndir*B,Cf,H,W]."""
    return _StemFn.apply(weight, bias, -1.0 if clamp_t is None else float(clamp_t), float(slope), *flows)
