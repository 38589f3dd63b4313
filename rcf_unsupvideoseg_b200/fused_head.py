"""The whole default head -- flow_feat_before_agg (conv 2->64 + LeakyReLU + conv 64->64 + LeakyReLU), masked pooling,
flow_feat_after_agg and the relaxed-common-fate loss (reference models/flow_aggregation_head_with_residual.py:84-101,
:235-310, :359-368) -- as ONE autograd node over the C ABI, with no ATen / cuDNN kernel on the path:

  forward   rcf_head_forward  = [residual up-sampling when predicted at a lower resolution]
                                 stem (mma.sync TF32; 3xTF32 in the fp32-grade mode) -> A1 as a bf16 (hi, lo) pair (fp32-grade),
                                 one fp16 tensor (TF32-class, torch's default) or one bf16 tensor (autocast) + sign bits
                                 weight pack (both orientations) -> tcgen05 conv -> pre-activation feature map (fp32 channels-last)
                                 rcf_forward: pooling (+ bias, LeakyReLU), segment MLP, loss
  backward  rcf_head_backward = rcf_backward: dM, dR, MLP gradients, conv-2 bias gradient, dfeat as a bf16 pair / scaled fp16 / bf16
                                 [gradient of the up-sampling] -> tcgen05 data gradient -> dA1
                                 tcgen05 weight gradient (K = pixels) -> dW2;  stem backward -> dW1, db1

One ctypes crossing and one device arena each way (the 96x96 / 48x48 training shapes are host-bound).
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
from .stem import _views


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


def _align(n: int) -> int:
    return (n + 255) & ~255


_AUX_BYTES = {}


def _aux_sizes(lib, dev, ndir, B, H, W, ks):
    """(wgrad workspace bytes, stem workspace bytes) for a shape, asked once."""
    key = (dev.index, ndir, B, H, W, ks)
    hit = _AUX_BYTES.get(key)
    if hit is None:
        a, b = C.c_size_t(), C.c_size_t()
        _lib.check(lib.rcf_conv64_wgrad_workspace_bytes(ndir * B, H, W, C.byref(a)), "rcf_conv64_wgrad_workspace_bytes")
        _lib.check(lib.rcf_stem_workspace_bytes(ndir, B, H, W, 64, ks, C.byref(b)), "rcf_stem_workspace_bytes")
        hit = _AUX_BYTES[key] = (a.value, b.value)
    return hit


class RcfHeadFn(torch.autograd.Function):
    """Inputs: spec (LossSpec with Cf = 64), nprod, stem_slope, masks [B,ndir,K,H,W], conv-1 weight/bias, conv-2 weight/bias,
    MLP w1,b1,w2,b2, then ndir flows [B,2,H,W] (no grad) and ndir residuals [B,2K,H,W].
    Returns (loss [ndir], total, *vis) like RcfMotionLossFn.

    One library call each way (rcf_head_forward / rcf_head_backward) and ONE device arena each way for everything that is
    not returned to autograd (activation pair, sign bits, packed weights, feature map, ctx, scratch): at the 96x96 / 48x48
    training shapes a step is bound by host time, not by the GPU."""

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
        nimg = ndir * B
        masks_v = masks if (masks.dtype == torch.float32 and _inner_dense(masks, 3)) else masks.float().contiguous()
        flows_v = _views(flows)
        # residuals predicted at a lower resolution (all directions alike): the library up-samples them itself
        rh, rw = resids[0].shape[-2:]
        lowres = (rh, rw) != (H, W)
        if lowres:
            assert all(tuple(r.shape) == (B, 2 * K, rh, rw) for r in resids)
            resids_v = [r.detach().float().contiguous() for r in resids]
        else:
            resids_v = [_as_dir_view(r) for r in resids]
        cw1c, cb1c, cw2c, cb2c, w1c, b1c, w2c, b2c = (t.detach().float().contiguous() for t in (cw1, cb1, cw2, cb2, w1, b1, w2, b2))
        ks = cw1c.shape[-1]

        desc = _make_desc(spec, B, ndir)
        desc.feat_nhwc = 1
        inp = _lib.RcfInputs()
        for i in range(ndir):
            inp.mask[i] = masks_v.data_ptr() + i * masks_v.stride(1) * 4
            desc.mask_bstride[i] = masks_v.stride(0)
            inp.flow[i] = flows_v[i].data_ptr(); desc.flow_bstride[i] = flows_v[i].stride(0)
            inp.resid[i] = resids_v[i].data_ptr()
            desc.resid_bstride[i] = 2 * K * P if lowres else resids_v[i].stride(0)
            desc.feat_bstride[i] = P * 64
        inp.w1, inp.b1, inp.w2, inp.b2 = (t.data_ptr() for t in (w1c, b1c, w2c, b2c))
        inp.feat_bias = cb2c.data_ptr()
        ctx_bytes, ws_bytes = _sizes(lib, desc, spec, B, ndir)

        # arena: [a_hi | a_lo? | sign | wpack x2 | feat | ctx | ws | up-sampled residuals?]
        act_b = nimg * P * 64 * 2
        offs, o = {}, 0
        for name, nbytes in (("a_hi", act_b), ("a_lo", act_b if nprod == 3 else 0), ("sign", nimg * P * 8),
                             ("wpack", 2 * c64.WPACK_BYTES), ("feat", nimg * P * 64 * 4), ("ctx", ctx_bytes), ("ws", ws_bytes),
                             ("resid_up", nimg * 2 * K * P * 4 if lowres else 0)):
            offs[name] = o
            o += _align(nbytes)
        arena = torch.empty(o, dtype=torch.uint8, device=dev)
        base = arena.data_ptr()
        hb = _lib.RcfHeadBuffers()
        hb.a_hi = base + offs["a_hi"]
        hb.a_lo = base + offs["a_lo"] if nprod == 3 else None
        hb.sign = base + offs["sign"]
        hb.wpack = base + offs["wpack"]
        hb.feat = base + offs["feat"]
        hb.resid_up = base + offs["resid_up"] if lowres else None

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
        with _lib.device_guard(dev):
            _lib.check(lib.rcf_head_forward(C.byref(desc), C.byref(inp), cw1c.data_ptr(), cb1c.data_ptr(), cw2c.data_ptr(), int(ks),
                                            float(stem_slope), int(nprod), int(rh) if lowres else 0, int(rw) if lowres else 0,
                                            C.byref(hb), loss_buf.data_ptr(), base + offs["ctx"],
                                            base + offs["ws"], C.byref(vis_struct) if vis_struct is not None else None,
                                            _lib.raw_stream(dev)), "rcf_head_forward")
        ctx.spec, ctx.ndir, ctx.B, ctx.nprod, ctx.stem_slope, ctx.ks = spec, ndir, B, nprod, float(stem_slope), int(ks)
        ctx.masks_shape = tuple(masks.shape)
        ctx.offs = offs
        ctx.resid_hw = (int(rh), int(rw)) if lowres else None
        ctx.save_for_backward(masks_v, arena, cw1c, cb2c, w1c, b1c, w2c, b2c, *flows_v, *([] if lowres else resids_v))
        ctx.mark_non_differentiable(*vis_tensors)
        ctx.set_materialize_grads(False)
        return (loss_buf[:ndir], loss_buf[ndir], *vis_tensors)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, grad_total, *grad_vis):
        lib = _lib.load_library()
        spec, ndir, B, nprod, ks = ctx.spec, ctx.ndir, ctx.B, ctx.nprod, ctx.ks
        n_in = 12 + 2 * ndir
        if grad_loss is None and grad_total is None:
            return (None,) * n_in
        K, H, W = spec.K, spec.H, spec.W
        P = H * W
        nimg = ndir * B
        saved = list(ctx.saved_tensors)
        masks_v, arena, cw1c, cb2c, w1c, b1c, w2c, b2c = saved[:8]
        lowres = ctx.resid_hw is not None
        flows_v, resids_v = saved[8:8 + ndir], (None if lowres else saved[8 + ndir:8 + 2 * ndir])
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
        for i in range(ndir):
            inp.mask[i] = masks_v.data_ptr() + i * masks_v.stride(1) * 4
            desc.mask_bstride[i] = masks_v.stride(0)
            inp.flow[i] = flows_v[i].data_ptr(); desc.flow_bstride[i] = flows_v[i].stride(0)
            if not lowres:
                inp.resid[i] = resids_v[i].data_ptr(); desc.resid_bstride[i] = resids_v[i].stride(0)
            desc.feat_bstride[i] = P * 64
            desc.dfeat_bstride[i] = P * 64
            if d_masks is not None:
                grads.dmask[i] = d_masks.data_ptr() + i * d_masks.stride(1) * 4
                desc.dmask_bstride[i] = d_masks.stride(0)
            if need_resid[i]:
                rshape = (B, 2 * K, *ctx.resid_hw) if lowres else (B, 2 * K, H, W)
                d_resids[i] = torch.empty(rshape, dtype=torch.float32, device=dev)
                grads.dresid[i] = d_resids[i].data_ptr(); desc.dresid_bstride[i] = d_resids[i].stride(0)
        d_cw1 = d_cb1 = d_cw2 = d_cb2 = None
        dmlp = [None] * 4
        if need_mlp:
            dmlp = [torch.empty(64, 64, 1, dtype=torch.float32, device=dev), torch.empty(64, dtype=torch.float32, device=dev),
                    torch.empty(2, 64, 1, dtype=torch.float32, device=dev), torch.empty(2, dtype=torch.float32, device=dev)]
            grads.dw1, grads.db1, grads.dw2, grads.db2 = (t.data_ptr() for t in dmlp)
        _, ws_bytes = _sizes(lib, desc, spec, B, ndir)

        base_f, offs = arena.data_ptr(), ctx.offs
        hb = _lib.RcfHeadBuffers()
        hb.a_hi = base_f + offs["a_hi"]
        hb.a_lo = base_f + offs["a_lo"] if nprod == 3 else None
        hb.sign = base_f + offs["sign"]
        hb.wpack = base_f + offs["wpack"]
        hb.feat = base_f + offs["feat"]
        hb.resid_up = base_f + offs["resid_up"] if lowres else None
        # backward arena: [ws | full-resolution dR? | g_hi | g_lo? | d_a1 | wgrad ws | stem ws]
        parts = [("ws", ws_bytes)]
        if lowres and any(need_resid):
            parts.append(("dresid_up", nimg * 2 * K * P * 4))
        if need_conv:
            wg_b, st_b = _aux_sizes(lib, dev, ndir, B, H, W, ks)
            act_b = nimg * P * 64 * 2
            parts += [("g_hi", act_b), ("g_lo", act_b if nprod == 3 else 0), ("d_a1", 2 * act_b), ("wgrad_ws", wg_b), ("stem_ws", st_b)]
        boffs, o = {}, 0
        for name, nbytes in parts:
            boffs[name] = o
            o += _align(nbytes)
        barena = torch.empty(max(o, 256), dtype=torch.uint8, device=dev)
        bbase = barena.data_ptr()
        if "dresid_up" in boffs:
            hb.dresid_up = bbase + boffs["dresid_up"]
        if need_conv:
            d_cw1 = torch.empty(tuple(cw1c.shape), dtype=torch.float32, device=dev)
            d_cb1 = torch.empty(64, dtype=torch.float32, device=dev)
            d_cw2 = torch.empty(64, 64, 3, 3, dtype=torch.float32, device=dev)
            d_cb2 = torch.empty(64, dtype=torch.float32, device=dev)
            grads.dfeat_bias = d_cb2.data_ptr()
            hb.g_hi = bbase + boffs["g_hi"]
            hb.g_lo = bbase + boffs["g_lo"] if nprod == 3 else None
            hb.d_a1 = bbase + boffs["d_a1"]
            hb.wgrad_ws = bbase + boffs["wgrad_ws"]
            hb.stem_ws = bbase + boffs["stem_ws"]
            hb.d_cw1, hb.d_cb1, hb.d_cw2 = d_cw1.data_ptr(), d_cb1.data_ptr(), d_cw2.data_ptr()
        if grad_loss is None:
            gl = grad_total.detach().to(torch.float32).contiguous()
            desc.grad_loss_total = 1
        else:
            gl = grad_loss.detach().to(torch.float32)
            if grad_total is not None:
                gl = gl + grad_total.detach().to(torch.float32)
            gl = gl.contiguous()
        with _lib.device_guard(dev):
            _lib.check(lib.rcf_head_backward(C.byref(desc), C.byref(inp), gl.data_ptr(), base_f + offs["ctx"], bbase + boffs["ws"],
                                             C.byref(grads), int(ks), float(ctx.stem_slope), int(nprod), int(need_conv),
                                             ctx.resid_hw[0] if lowres else 0, ctx.resid_hw[1] if lowres else 0,
                                             C.byref(hb), _lib.raw_stream(dev)), "rcf_head_backward")
        return (None, None, None, d_masks, d_cw1, d_cb1, d_cw2, d_cb2, *dmlp, *([None] * ndir), *d_resids)
