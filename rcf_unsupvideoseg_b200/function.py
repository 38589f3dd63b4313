"""torch.autograd.Function over the C-ABI library: the RCF motion loss, forward + backward.

PyTorch is used here for device memory, streams and the autograd hook only; every per-pixel and
per-segment computation of reference models/flow_aggregation_head_with_residual.py:235-310 and
:359-368 runs inside librcf_loss.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


@dataclass(frozen=True)
class LossSpec:
    """Static description of one loss call (mirrors the reference constructor flags that reach the math)."""
    K: int
    H: int
    W: int
    D: int = 0                      # 0 free_residual, 2 affine, 5 quadratic
    Cf: int = 0                     # pooled feature channels (0 => theta supplied)
    robust: bool = False
    eps: float = 0.01
    q: float = 0.4
    resid_scale: float = 10.0
    pred_div: float = 10.0
    clamp_t: Optional[float] = None
    unbounded_residual: bool = False
    inv_n: float = 0.0              # 0 => 1/(B*2*H*W); set for batch-sharded (multi-GPU) use
    want_vis: bool = False
    vis_scale: Tuple[float, float] = (1.0, 1.0)
    feat_lrelu_slope: float = 1.0   # 0.1: `feat` is the pre-activation of the last conv, LeakyReLU fused into the kernels

    @property
    def theta_mode(self) -> int:
        return 1 if self.Cf > 0 else 0


def _inner_dense(t: torch.Tensor, nd: int) -> bool:
    """last `nd` dims laid out densely (row-major) -- what the kernels require of [C,H,W] blocks."""
    exp = 1
    for size, stride in zip(reversed(t.shape[-nd:]), reversed(t.stride()[-nd:])):
        if size != 1 and stride != exp:
            return False
        exp *= size
    return True


def _as_dir_view(t: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] fp32 CUDA tensor whose [C,H,W] block is dense; batch stride is free."""
    if t.dtype != torch.float32:
        t = t.float()
    if not _inner_dense(t, 3):
        t = t.contiguous()
    return t


def _feat_layout(feat: torch.Tensor, Cf: int, H: int, W: int):
    """feat [ndir,B,Cf,H,W] fp32: returns (tensor, nhwc) with either dense [Cf,H,W] planes or dense channels-last
    frames (what cuDNN returns for channels_last inputs); anything else is made plane-contiguous (one copy)."""
    if feat.dtype != torch.float32:
        feat = feat.float()
    st = feat.stride()
    if Cf % 4 == 0 and Cf <= 128 and 256 % (Cf // 4) == 0 and st[2] == 1 and st[4] == Cf and st[3] == W * Cf \
            and st[1] % 4 == 0 and st[0] % 4 == 0 and feat.data_ptr() % 16 == 0:
        return feat, True
    if _inner_dense(feat, 3):
        return feat, False
    return feat.contiguous(), False


_SIZE_CACHE = {}


def _sizes(lib, desc, spec, B, ndir):
    key = (spec.K, spec.H, spec.W, spec.D, spec.Cf, B, ndir, int(desc.feat_nhwc))
    hit = _SIZE_CACHE.get(key)
    if hit is None:
        ctx_bytes, ws_bytes = C.c_size_t(), C.c_size_t()
        _lib.check(lib.rcf_query_sizes(C.byref(desc), C.byref(ctx_bytes), C.byref(ws_bytes)), "rcf_query_sizes")
        hit = _SIZE_CACHE[key] = (ctx_bytes.value, ws_bytes.value)
    return hit


def _make_desc(spec: LossSpec, B: int, ndir: int) -> _lib.RcfDesc:
    d = _lib.RcfDesc()
    d.B, d.K, d.H, d.W, d.Cf, d.D, d.ndir = B, spec.K, spec.H, spec.W, spec.Cf, spec.D, ndir
    d.theta_mode = spec.theta_mode
    d.robust = int(spec.robust)
    d.unbounded_residual = int(spec.unbounded_residual)
    d.eps, d.q = spec.eps, spec.q
    d.resid_scale, d.pred_div = spec.resid_scale, spec.pred_div
    d.clamp_t = -1.0 if spec.clamp_t is None else float(spec.clamp_t)
    d.inv_n = spec.inv_n
    d.feat_lrelu_slope = spec.feat_lrelu_slope
    return d


class RcfMotionLossFn(torch.autograd.Function):
    """loss[dir] = mean(phi(F_dir - pred_dir)) for ndir directions.

    Inputs (all fp32 CUDA):
      masks  [B, ndir, K, H, W]           (grad)
      feat   [ndir, B, Cf, H, W]          (grad)  + w1,b1,w2,b2 (grad)      when spec.Cf > 0
             (direction-major, i.e. the conv output of the concatenated [fw; bw] flow batch, viewed 5-D)
      flows  tuple of ndir  [B, 2, H, W]  (no grad; the clamp is applied inside the kernels)
      resids tuple of ndir  [B, 2K, H, W] (grad)
      thetas tuple of ndir  [B, 2, K]     (grad)                            when spec.Cf == 0
    Returns loss [ndir], total (0-dim, = sum of the directions, reference :397) and, when spec.want_vis, the
    un-differentiable tensors (gt, pred, agg, res[, aff]) each [B, 2*ndir, H, W].  `loss` and `total` are two views of
    one (ndir+1)-float buffer the library fills: differentiating `total` (what the reference's caller does,
    rcf_model.py:464-470) hands its 0-dim gradient straight to the library, with no select/add/fill kernels of
    autograd in between.
    """

    @staticmethod
    def forward(ctx, spec: LossSpec, masks, feat, feat_bias, w1, b1, w2, b2, *per_dir):
        lib = _lib.load_library()
        ndir = masks.shape[1]
        assert len(per_dir) == 3 * ndir
        flows = per_dir[0:ndir]
        resids = per_dir[ndir:2 * ndir]
        thetas = per_dir[2 * ndir:3 * ndir]
        if not masks.is_cuda:
            raise RuntimeError("RcfMotionLossFn needs CUDA tensors: the loss has no CPU implementation")
        B, _, K, H, W = masks.shape
        assert (K, H, W) == (spec.K, spec.H, spec.W), f"mask shape {tuple(masks.shape)} vs spec {spec}"
        dev = masks.device
        P = H * W

        masks_v = masks if (masks.dtype == torch.float32 and _inner_dense(masks, 3)) else masks.float().contiguous()
        flows_v = [_as_dir_view(f) for f in flows]
        resids_v = [_as_dir_view(r) for r in resids]
        feat_v, nhwc = None, False
        if spec.Cf > 0:
            assert feat is not None and tuple(feat.shape) == (ndir, B, spec.Cf, H, W), \
                f"feature map shape {None if feat is None else tuple(feat.shape)}"
            feat_v, nhwc = _feat_layout(feat, spec.Cf, H, W)
        fb = None
        if feat_bias is not None:
            if not nhwc:
                raise RuntimeError("feat_bias needs a channels-last feature map (fuse the bias into the conv otherwise)")
            fb = feat_bias.detach().float().contiguous()
            assert fb.numel() == spec.Cf
        thetas_v = [t.float().contiguous() if t is not None else None for t in thetas]
        for f in flows_v:
            assert f.shape == (B, 2, H, W), f"flow shape {tuple(f.shape)}"
        for r in resids_v:
            assert r.shape == (B, 2 * K, H, W), f"residual shape {tuple(r.shape)}"
        if spec.Cf > 0:
            w1c, b1c, w2c, b2c = (t.detach().float().contiguous() for t in (w1, b1, w2, b2))
            assert w1c.numel() == spec.Cf * spec.Cf and w2c.numel() == 2 * spec.Cf
        else:
            for t in thetas_v:
                assert t is not None and t.shape == (B, 2, K), "theta shape"
            w1c = b1c = w2c = b2c = None

        desc = _make_desc(spec, B, ndir)
        desc.feat_nhwc = int(nhwc)
        inp = _lib.RcfInputs()
        for i in range(ndir):
            inp.mask[i] = masks_v.data_ptr() + i * masks_v.stride(1) * 4
            desc.mask_bstride[i] = masks_v.stride(0)
            inp.flow[i] = flows_v[i].data_ptr(); desc.flow_bstride[i] = flows_v[i].stride(0)
            inp.resid[i] = resids_v[i].data_ptr(); desc.resid_bstride[i] = resids_v[i].stride(0)
            if spec.Cf > 0:
                inp.feat[i] = feat_v.data_ptr() + i * feat_v.stride(0) * 4; desc.feat_bstride[i] = feat_v.stride(1)
            else:
                inp.theta[i] = thetas_v[i].data_ptr()
        if spec.Cf > 0:
            inp.w1, inp.b1, inp.w2, inp.b2 = (t.data_ptr() for t in (w1c, b1c, w2c, b2c))
            if fb is not None:
                inp.feat_bias = fb.data_ptr()

        ctx_bytes, ws_bytes = _sizes(lib, desc, spec, B, ndir)
        ctx_buf = torch.empty(ctx_bytes, dtype=torch.uint8, device=dev)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss_buf = torch.empty(ndir + 1, dtype=torch.float32, device=dev)

        vis_tensors: Tuple[torch.Tensor, ...] = ()
        vis_struct = None
        if spec.want_vis:
            n_out = 5 if spec.D > 0 else 4
            vis_tensors = tuple(torch.empty(B, 2 * ndir, H, W, dtype=torch.float32, device=dev) for _ in range(n_out))
            vis_struct = _lib.RcfVisOut()
            vis_struct.gt, vis_struct.pred, vis_struct.agg, vis_struct.res = (t.data_ptr() for t in vis_tensors[:4])
            vis_struct.aff = vis_tensors[4].data_ptr() if n_out == 5 else None
            desc.vis_bstride = 2 * ndir * P
            desc.vis_dstride = 2 * P
            desc.vis_scale[0], desc.vis_scale[1] = spec.vis_scale

        stream = _lib.raw_stream(dev)
        with _lib.device_guard(dev):
            _lib.check(lib.rcf_forward(C.byref(desc), C.byref(inp), loss_buf.data_ptr(), ctx_buf.data_ptr(), ws.data_ptr(),
                                       C.byref(vis_struct) if vis_struct is not None else None, stream), "rcf_forward")

        ctx.spec, ctx.ndir, ctx.B, ctx.nhwc = spec, ndir, B, nhwc
        ctx.has_fb = fb is not None
        ctx.masks_shape = tuple(masks.shape)
        ctx.save_for_backward(masks_v, ctx_buf, *flows_v, *resids_v,
                              *([feat_v] if feat_v is not None else []), *[t for t in thetas_v if t is not None],
                              *([w1c, b1c, w2c, b2c] if spec.Cf > 0 else []), *([fb] if fb is not None else []))
        ctx.mark_non_differentiable(*vis_tensors)
        ctx.set_materialize_grads(False)     # no zero-filled gradients for the visualisation outputs / unused loss views
        return (loss_buf[:ndir], loss_buf[ndir], *vis_tensors)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loss, grad_total, *grad_vis):
        lib = _lib.load_library()
        spec, ndir, B = ctx.spec, ctx.ndir, ctx.B
        n_in = 8 + 3 * ndir
        if grad_loss is None and grad_total is None:
            return (None,) * n_in
        K, H, W, Cf = spec.K, spec.H, spec.W, spec.Cf
        saved = list(ctx.saved_tensors)
        masks_v, ctx_buf = saved[0], saved[1]
        flows_v = saved[2:2 + ndir]
        resids_v = saved[2 + ndir:2 + 2 * ndir]
        rest = saved[2 + 2 * ndir:]
        dev = masks_v.device
        # needs_input_grad: (spec, masks, feat, feat_bias, w1, b1, w2, b2, *flows, *resids, *thetas)
        need = ctx.needs_input_grad
        need_masks = need[1]
        need_feat = need[2]
        need_fb = need[3] and ctx.has_fb
        need_w = any(need[4:8])
        need_resid = [need[8 + ndir + i] for i in range(ndir)]
        need_theta = [need[8 + 2 * ndir + i] for i in range(ndir)]

        desc = _make_desc(spec, B, ndir)
        desc.feat_nhwc = int(ctx.nhwc)
        inp = _lib.RcfInputs()
        grads = _lib.RcfGrads()
        d_feat = None
        if Cf > 0:
            feat_v = rest[0]
            w1c, b1c, w2c, b2c = rest[1:5]
            inp.w1, inp.b1, inp.w2, inp.b2 = (t.data_ptr() for t in (w1c, b1c, w2c, b2c))
            if ctx.has_fb:
                inp.feat_bias = rest[5].data_ptr()
            if need_feat:
                if ctx.nhwc:    # gradient in the layout of the feature map (cuDNN's backward then needs no transposes either)
                    d_feat = torch.empty((ndir * B, Cf, H, W), dtype=torch.float32, device=dev,
                                         memory_format=torch.channels_last).view(ndir, B, Cf, H, W)
                else:
                    d_feat = torch.empty(ndir, B, Cf, H, W, dtype=torch.float32, device=dev)
        else:
            thetas_v = rest[0:ndir]

        d_masks = torch.empty(ctx.masks_shape, dtype=torch.float32, device=dev) if need_masks else None
        d_resids, d_thetas = [None] * ndir, [None] * ndir
        for i in range(ndir):
            inp.mask[i] = masks_v.data_ptr() + i * masks_v.stride(1) * 4
            desc.mask_bstride[i] = masks_v.stride(0)
            inp.flow[i] = flows_v[i].data_ptr(); desc.flow_bstride[i] = flows_v[i].stride(0)
            inp.resid[i] = resids_v[i].data_ptr(); desc.resid_bstride[i] = resids_v[i].stride(0)
            if d_masks is not None:
                grads.dmask[i] = d_masks.data_ptr() + i * d_masks.stride(1) * 4
                desc.dmask_bstride[i] = d_masks.stride(0)
            if need_resid[i]:
                d_resids[i] = torch.empty(B, 2 * K, H, W, dtype=torch.float32, device=dev)
                grads.dresid[i] = d_resids[i].data_ptr(); desc.dresid_bstride[i] = d_resids[i].stride(0)
            if Cf > 0:
                inp.feat[i] = feat_v.data_ptr() + i * feat_v.stride(0) * 4; desc.feat_bstride[i] = feat_v.stride(1)
                if d_feat is not None:
                    grads.dfeat[i] = d_feat.data_ptr() + i * d_feat.stride(0) * 4; desc.dfeat_bstride[i] = d_feat.stride(1)
            else:
                inp.theta[i] = thetas_v[i].data_ptr()
                if need_theta[i]:
                    d_thetas[i] = torch.empty(B, 2, K, dtype=torch.float32, device=dev)
                    grads.dtheta[i] = d_thetas[i].data_ptr()
        d_fb = None
        if Cf > 0 and need_fb:
            d_fb = torch.empty(Cf, dtype=torch.float32, device=dev)
            grads.dfeat_bias = d_fb.data_ptr()
        dw = [None, None, None, None]
        if Cf > 0 and need_w:
            dw = [torch.empty(Cf, Cf, 1, dtype=torch.float32, device=dev), torch.empty(Cf, dtype=torch.float32, device=dev),
                  torch.empty(2, Cf, 1, dtype=torch.float32, device=dev), torch.empty(2, dtype=torch.float32, device=dev)]
            grads.dw1, grads.db1, grads.dw2, grads.db2 = (t.data_ptr() for t in dw)

        _, ws_bytes = _sizes(lib, desc, spec, B, ndir)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        if grad_loss is None:            # only the total was differentiated: its 0-dim gradient serves every direction
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
        return (None, d_masks, d_feat, d_fb, *dw, *([None] * ndir), *d_resids, *d_thetas)


def rcf_motion_loss(spec: LossSpec, masks: torch.Tensor, flows: Sequence[torch.Tensor],
                    resids: Sequence[torch.Tensor], *, feats: Optional[Sequence[torch.Tensor]] = None,
                    mlp: Optional[Sequence[torch.Tensor]] = None, thetas: Optional[Sequence[torch.Tensor]] = None,
                    feat_bias: Optional[torch.Tensor] = None, with_total: bool = False):
    """Functional entry point.  masks [B,ndir,K,H,W]; flows/resids (and feats or thetas) per direction.

    Returns (loss [ndir], vis tuple); with `with_total=True`, (loss [ndir], total 0-dim, vis tuple).  Exactly one of (feats + mlp weights) or thetas must be given,
    consistently with spec.Cf.  `feats` is either the direction-major 5-D tensor [ndir,B,Cf,H,W] (preferred:
    one conv call over the concatenated directions, no copies) or a sequence of per-direction [B,Cf,H,W] tensors
    (stacked here, which costs a copy).  `feat_bias` [Cf] (channels-last feats only) is the bias of the conv that produced
    `feats`, applied inside the kernels so that the conv itself can run bias-free.
    """
    ndir = masks.shape[1]
    feat = None
    if spec.Cf > 0:
        assert feats is not None and mlp is not None and len(mlp) == 4
        w1, b1, w2, b2 = mlp
        feat = feats if torch.is_tensor(feats) else torch.stack(list(feats), 0)
        per_dir = (*flows, *resids, *([None] * ndir))
    else:
        assert thetas is not None
        w1 = b1 = w2 = b2 = None
        per_dir = (*flows, *resids, *thetas)
    out = RcfMotionLossFn.apply(spec, masks, feat, feat_bias if feat is not None else None, w1, b1, w2, b2, *per_dir)
    if with_total:
        return out[0], out[1], tuple(out[2:])
    return out[0], tuple(out[2:])
