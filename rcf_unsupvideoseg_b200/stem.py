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
                                        torch.cuda.current_stream(dev).cuda_stream), "rcf_stem_forward")
    return act, sign


def stem_backward_raw(fl, w_shape, clamp_t: float, slope: float, act, sign, dact):
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
                                         db.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
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


_SIDE = {}


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    s = _SIDE.get(idx)
    if s is None:
        s = _SIDE[idx] = torch.cuda.Stream(device=idx)
    return s


class _StemConvFn(torch.autograd.Function):
    """stem (this library) + the bias-free second convolution (cuDNN) as ONE autograd node, so that the backward can be
    scheduled: the second conv's weight gradient (independent: it only needs dG and the stem output) runs on a side stream
    while its data gradient and, behind it, the stem's weight/bias gradient run on the main stream; the streams join before
    the four parameter gradients are handed back.  At the 96x96 / 48x48 training shapes neither chain fills the GPU
    (wgrad 52 us; dgrad + stem backward 57 us at 8x2x96x96) and they overlap; the fork/join is expressed with stream waits,
    so a captured CUDA graph keeps the two branches parallel."""

    @staticmethod
    def forward(ctx, w1, b1, w2, clamp_t: float, slope: float, conv_conf, *flows):
        fl = _views(flows)
        w = w1.detach().float().contiguous()
        b = b1.detach().float().contiguous()
        act, sign = stem_forward_raw(fl, w, b, clamp_t, slope)
        stride, padding, dilation, groups = conv_conf
        w2d = w2.detach()
        feat = torch.ops.aten.convolution(act, w2d, None, stride, padding, dilation, False, [0, 0], groups)
        ctx.conf = (tuple(w.shape), float(clamp_t), float(slope), len(fl), conv_conf)
        ctx.has_sign = sign is not None
        ctx.save_for_backward(act, w2d, *([sign] if sign is not None else []), *fl)
        return feat

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        w_shape, clamp_t, slope, ndir, (stride, padding, dilation, groups) = ctx.conf
        saved = list(ctx.saved_tensors)
        act, w2d = saved[0], saved[1]
        sign = saved[2] if ctx.has_sign else None
        fl = saved[3:] if ctx.has_sign else saved[2:]
        need_stem = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_w2 = ctx.needs_input_grad[2]
        g = g.contiguous(memory_format=torch.channels_last)
        main = torch.cuda.current_stream(g.device)
        gw2 = dw = db = None
        if need_w2 and need_stem:
            side = _side_stream(g.device)
            side.wait_stream(main)                     # dG is ready
            with torch.cuda.stream(side):
                gw2 = torch.ops.aten.convolution_backward(g, act, w2d, None, stride, padding, dilation, False, [0, 0], groups,
                                                          [False, True, False])[1]
            dact = torch.ops.aten.convolution_backward(g, act, w2d, None, stride, padding, dilation, False, [0, 0], groups,
                                                       [True, False, False])[0]
            dw, db = stem_backward_raw(fl, w_shape, clamp_t, slope, act, sign, dact)
            main.wait_stream(side)                     # join: everything returned below is complete w.r.t. `main`
            gw2.record_stream(main)
            g.record_stream(side)
            act.record_stream(side)
        else:
            res = torch.ops.aten.convolution_backward(g, act, w2d, None, stride, padding, dilation, False, [0, 0], groups,
                                                      [bool(need_stem), bool(need_w2), False])
            gw2 = res[1]
            if need_stem:
                dw, db = stem_backward_raw(fl, w_shape, clamp_t, slope, act, sign, res[0])
        return (dw, db, gw2, None, None, None, *([None] * ndir))


def flow_stem(flows: Sequence[torch.Tensor], weight: torch.Tensor, bias: torch.Tensor, clamp_t: Optional[float],
              slope: float = 0.1) -> torch.Tensor:
    """LeakyReLU(conv(clamp(flow)) + bias) for 1 or 2 directions of [B,2,H,W] flows -> channels-last [ndir*B,Cf,H,W]."""
    return _StemFn.apply(weight, bias, -1.0 if clamp_t is None else float(clamp_t), float(slope), *flows)


def flow_stem_conv(flows: Sequence[torch.Tensor], w1: torch.Tensor, b1: torch.Tensor, clamp_t: Optional[float], slope: float,
                   w2: torch.Tensor, stride, padding, dilation, groups) -> torch.Tensor:
    """conv2(LeakyReLU(conv1(clamp(flow)) + b1)) without conv2's bias (the pooling kernels add it), channels-last; one
    autograd node whose backward overlaps conv2's weight gradient with its data gradient + the stem's gradients."""
    conf = (tuple(stride), tuple(padding), tuple(dilation), int(groups))
    return _StemConvFn.apply(w1, b1, w2, -1.0 if clamp_t is None else float(clamp_t), float(slope), conf, *flows)
