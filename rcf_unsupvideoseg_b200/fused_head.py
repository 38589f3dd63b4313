"""The whole default head -- flow_feat_before_agg (conv 2->64 + LeakyReLU + conv 64->64 + LeakyReLU), masked pooling,
flow_feat_after_agg and the relaxed-common-fate loss (reference models/flow_aggregation_head_with_residual.py:84-101,
:235-310, :359-368) -- as ONE autograd node over the C ABI, with no ATen / cuDNN kernel on the path:

  forward   rcf_stem_forward_bf16 (mma.sync 3xTF32)  ->  A1 as a bf16 (hi, lo) pair + sign bits
            rcf_conv64_forward   (TMA + tcgen05.mma)  ->  pre-activation feature map, fp32 channels-last
            rcf_forward                               ->  pooling (+ bias, LeakyReLU), segment MLP, loss
  backward  rcf_backward                              ->  dM, dR, MLP gradients, conv-2 bias gradient, dfeat as a bf16 pair
            rcf_conv64_forward (transposed weights)   ->  dA1
            rcf_conv64_wgrad   (tcgen05, K = pixels)  ->  dW2
            rcf_stem_backward                         ->  dW1, db1

The intermediate activations never pass through autograd, so they can travel in the operand format of the tensor-core
kernels (two bf16 tensors, x ~ hi + lo) instead of fp32.  Used by FlowAggregationHeadWithResidual when the head has its
default shape (64 channels, 3x3 kernels); other shapes take the general path in head.py / function.py.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import conv64 as c64
from .function import LossSpec, _as_dir_view, _inner_dense, _make_desc, _sizes
from .stem import _views, stem_backward_raw


def stem_forward_pair(fl, w, b, clamp_t: float, slope: float, want_lo: bool, nprod: int = 3):
    """First conv + LeakyReLU; returns (act_hi, act_lo or None, sign): bf16 channels-last [ndir*B,64,H,W] and the sign bits.
    nprod >= 3: 3xTF32 (fp32-grade); below: one TF32 product, like cuDNN under allow_tf32."""
    lib = _lib.load_library()
    ndir = len(fl)
    B, _, H, W = fl[0].shape
    dev = fl[0].device
    ks = w.shape[-1]
    hi = torch.empty((ndir * B, 64, H, W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
    lo = torch.empty_like(hi) if want_lo else None
    sign = torch.empty(ndir * B * H * W * 2, dtype=torch.int32, device=dev)
    ptrs = (C.c_void_p * 2)(*[f.data_ptr() for f in fl], *([None] * (2 - ndir)))
    strides = (C.c_int64 * 2)(*[f.stride(0) for f in fl], *([0] * (2 - ndir)))
    with _lib.device_guard(dev):
        _lib.check(lib.rcf_stem_forward_bf16(ptrs, strides, ndir, B, H, W, ks, w.data_ptr(), b.data_ptr(), float(clamp_t),
                                             float(slope), hi.data_ptr(), lo.data_ptr() if lo is not None else None,
                                             sign.data_ptr(), int(nprod), _lib.raw_stream(dev)),
                   "rcf_stem_forward_bf16")
    return hi, lo, sign


class RcfHeadFn(torch.autograd.Function):
    """Inputs: spec (LossSpec with Cf = 64), nprod, stem_slope, masks [B,ndir,K,H,W], conv-1 weight/bias, conv-2 weight/bias,
    MLP w1,b1,w2,b2, then ndir flows [B,2,H,W] (no grad) and ndir residuals [B,2K,H,W].
    Returns (loss [ndir], total, *vis) like RcfMotionLossFn."""

    @staticmethod
    def forward(ctx, spec: LossSpec, nprod: int, stem_slope: float, masks, cw1, cb1, cw2, cb2, w1, b1, w2, b2, *per_dir):
        lib = _lib.load_library()
        ndir = masks.shape[1]
        flows, resids = per_dir[:ndir], per_dir[ndir:2 * ndir]
        if not masks.is_cuda:
            raise RuntimeError("RcfHeadFn needs CUDA tensors: the head has no CPU implementation")
        B, _, K, H, W = masks.shape
        assert spec.Cf == 64 and (K, H, W) == (spec.K, spec.H, spec.W)
        dev = masks.device
        P = H * W
        masks_v = masks if (masks.dtype == torch.float32 and _inner_dense(masks, 3)) else masks.float().contiguous()
        flows_v = _views(flows)
        resids_v = [_as_dir_view(r) for r in resids]
        cw1c, cb1c, cw2c, cb2c, w1c, b1c, w2c, b2c = (t.detach().float().contiguous() for t in (cw1, cb1, cw2, cb2, w1, b1, w2, b2))

        # conv branch: stem (bf16 pair out) -> tcgen05 conv (fp32 pre-activation, channels-last, bias applied by the pooling kernels)
        clamp = -1.0 if spec.clamp_t is None else float(spec.clamp_t)
        a_hi, a_lo, sign = stem_forward_pair(flows_v, cw1c, cb1c, clamp, stem_slope, want_lo=(nprod == 3), nprod=nprod)
        need_conv_grad = any(ctx.needs_input_grad[4:8])
        if need_conv_grad:          # the data-gradient operator of the same weights is packed by the same launch
            wp_fwd, wp_bwd = c64.pack_weights_both(cw2c)
        else:
            wp_fwd, wp_bwd = c64.pack_weights(cw2c, False), None
        feat = c64.conv64_pair(a_hi, a_lo, wp_fwd, nprod)          # [ndir*B,64,H,W] channels-last

        desc = _make_desc(spec, B, ndir)
        desc.feat_nhwc = 1
        inp = _lib.RcfInputs()
        for i in range(ndir):
            inp.mask[i] = masks_v.data_ptr() + i * masks_v.stride(1) * 4
            desc.mask_bstride[i] = masks_v.stride(0)
            inp.flow[i] = flows_v[i].data_ptr(); desc.flow_bstride[i] = flows_v[i].stride(0)
            inp.resid[i] = resids_v[i].data_ptr(); desc.resid_bstride[i] = resids_v[i].stride(0)
            inp.feat[i] = feat.data_ptr() + i * B * P * 64 * 4; desc.feat_bstride[i] = P * 64
        inp.w1, inp.b1, inp.w2, inp.b2 = (t.data_ptr() for t in (w1c, b1c, w2c, b2c))
        inp.feat_bias = cb2c.data_ptr()
        ctx_bytes, ws_bytes = _sizes(lib, desc, spec, B, ndir)
        ctx_buf = torch.empty(ctx_bytes, dtype=torch.uint8, device=dev)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss_buf = torch.empty(ndir + 1, dtype=torch.float32, device=dev)
        vis_tensors, vis_struct = (), None
        if spec.want_vis:
            n_out = 5 if spec.D > 0 else 4
            vis_tensors = tuple(torch.empty(B, 2 * ndir, H, W, dtype=torch.float32, device=dev) for _ in range(n_out))
            vis_struct = _lib.RcfVisOut()
            vis_struct.gt, vis_struct.pred, vis_struct.agg, vis_struct.res = (t.data_ptr() for t in vis_tensors[:4])
            vis_struct.aff = vis_tensors[4].data_ptr() if n_out == 5 else None
            desc.vis_bstride, desc.vis_dstride = 2 * ndir * P, 2 * P
            desc.vis_scale[0], desc.vis_scale[1] = spec.vis_scale
        stream = _lib.raw_stream(dev)
        with _lib.device_guard(dev):
            _lib.check(lib.rcf_forward(C.byref(desc), C.byref(inp), loss_buf.data_ptr(), ctx_buf.data_ptr(), ws.data_ptr(),
                                       C.byref(vis_struct) if vis_struct is not None else None, stream), "rcf_forward")
        ctx.spec, ctx.ndir, ctx.B, ctx.nprod, ctx.stem_slope, ctx.clamp = spec, ndir, B, nprod, stem_slope, clamp
        ctx.masks_shape = tuple(masks.shape)
        ctx.has_lo = a_lo is not None
        ctx.has_wp = wp_bwd is not None
        ctx.save_for_backward(masks_v, ctx_buf, feat, a_hi, sign, cw1c, cw2c, cb2c, w1c, b1c, w2c, b2c,
                              *([a_lo] if a_lo is not None else []), *([wp_bwd] if wp_bwd is not None else []),
                              *flows_v, *resids_v)
        ctx.mark_non_differentiable(*vis_tensors)
        ctx.set_materialize_grads(False)
        return (loss_buf[:ndir], loss_buf[ndir], *vis_tensors)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, grad_total, *grad_vis):
        lib = _lib.load_library()
        spec, ndir, B, nprod = ctx.spec, ctx.ndir, ctx.B, ctx.nprod
        n_in = 12 + 2 * ndir
        if grad_loss is None and grad_total is None:
            return (None,) * n_in
        K, H, W = spec.K, spec.H, spec.W
        P = H * W
        saved = list(ctx.saved_tensors)
        masks_v, ctx_buf, feat, a_hi, sign, cw1c, cw2c, cb2c, w1c, b1c, w2c, b2c = saved[:12]
        rest = saved[12:]
        a_lo = rest.pop(0) if ctx.has_lo else None
        wp_bwd = rest.pop(0) if ctx.has_wp else None
        flows_v, resids_v = rest[:ndir], rest[ndir:2 * ndir]
        dev = masks_v.device
        need = ctx.needs_input_grad          # (spec, nprod, slope, masks, cw1, cb1, cw2, cb2, w1, b1, w2, b2, *flows, *resids)
        need_masks = need[3]
        need_conv = any(need[4:8])
        need_mlp = any(need[8:12])
        need_resid = [need[12 + ndir + i] for i in range(ndir)]

        desc = _make_desc(spec, B, ndir)
        desc.feat_nhwc = 1
        inp = _lib.RcfInputs()
        grads = _lib.RcfGrads()
        inp.w1, inp.b1, inp.w2, inp.b2 = (t.data_ptr() for t in (w1c, b1c, w2c, b2c))
        inp.feat_bias = cb2c.data_ptr()
        d_masks = torch.empty(ctx.masks_shape, dtype=torch.float32, device=dev) if need_masks else None
        d_resids = [None] * ndir
        g_hi = g_lo = None
        if need_conv:
            g_hi = torch.empty((ndir * B, 64, H, W), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last)
            g_lo = torch.empty_like(g_hi) if nprod >= 2 else None
        for i in range(ndir):
            inp.mask[i] = masks_v.data_ptr() + i * masks_v.stride(1) * 4
            desc.mask_bstride[i] = masks_v.stride(0)
            inp.flow[i] = flows_v[i].data_ptr(); desc.flow_bstride[i] = flows_v[i].stride(0)
            inp.resid[i] = resids_v[i].data_ptr(); desc.resid_bstride[i] = resids_v[i].stride(0)
            inp.feat[i] = feat.data_ptr() + i * B * P * 64 * 4; desc.feat_bstride[i] = P * 64
            if d_masks is not None:
                grads.dmask[i] = d_masks.data_ptr() + i * d_masks.stride(1) * 4
                desc.dmask_bstride[i] = d_masks.stride(0)
            if need_resid[i]:
                d_resids[i] = torch.empty(B, 2 * K, H, W, dtype=torch.float32, device=dev)
                grads.dresid[i] = d_resids[i].data_ptr(); desc.dresid_bstride[i] = d_resids[i].stride(0)
            if g_hi is not None:
                grads.dfeat_hi[i] = g_hi.data_ptr() + i * B * P * 64 * 2
                if g_lo is not None:
                    grads.dfeat_lo[i] = g_lo.data_ptr() + i * B * P * 64 * 2
                desc.dfeat_bstride[i] = P * 64
        d_cb2 = None
        if need_conv:
            d_cb2 = torch.empty(64, dtype=torch.float32, device=dev)
            grads.dfeat_bias = d_cb2.data_ptr()
        dmlp = [None] * 4
        if need_mlp:
            dmlp = [torch.empty(64, 64, 1, dtype=torch.float32, device=dev), torch.empty(64, dtype=torch.float32, device=dev),
                    torch.empty(2, 64, 1, dtype=torch.float32, device=dev), torch.empty(2, dtype=torch.float32, device=dev)]
            grads.dw1, grads.db1, grads.dw2, grads.db2 = (t.data_ptr() for t in dmlp)
        _, ws_bytes = _sizes(lib, desc, spec, B, ndir)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        if grad_loss is None:
            gl = grad_total.detach().to(torch.float32).contiguous()
            desc.grad_loss_total = 1
        else:
            gl = grad_loss.detach().to(torch.float32)
            if grad_total is not None:
                gl = gl + grad_total.detach().to(torch.float32)
            gl = gl.contiguous()
        stream = _lib.raw_stream(dev)
        with _lib.device_guard(dev):
            _lib.check(lib.rcf_backward(C.byref(desc), C.byref(inp), gl.data_ptr(), ctx_buf.data_ptr(), ws.data_ptr(),
                                        C.byref(grads), stream), "rcf_backward")
        d_cw1 = d_cb1 = d_cw2 = None
        if need_conv:
            if wp_bwd is None:
                wp_bwd = c64.pack_weights(cw2c, True)
            d_a1 = c64.conv64_pair(g_hi, g_lo if nprod == 3 else None, wp_bwd, nprod)     # data gradient
            d_cw2 = c64.conv64_wgrad_pair(a_hi, a_lo, g_hi, g_lo, nprod)
            d_cw1, d_cb1 = stem_backward_raw(flows_v, tuple(cw1c.shape), ctx.clamp, ctx.stem_slope, None, sign, d_a1, nprod=ctx.nprod)
        return (None, None, None, d_masks, d_cw1, d_cb1, d_cw2, d_cb2, *dmlp, *([None] * ndir), *d_resids)
