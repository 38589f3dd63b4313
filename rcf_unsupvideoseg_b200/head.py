"""Drop-in replacement of the reference's FlowAggregationHeadWithResidual.

Mirrors /root/reference/models/flow_aggregation_head_with_residual.py:33-399: same class name,
constructor keywords and defaults, parameter names (``flow_feat_before_agg.{0,2}``,
``flow_feat_after_agg.{0,2}`` -- checkpoints load unchanged), ``forward`` signature and the
``(flows, flow_loss)`` return structure.  The reference instantiates heads by class name from the
module globals of models/rcf_model.py (:75); see INTEGRATION.md for the one-line rebinding.

What runs where:
  * default head shape (64 feature channels, 3x3 kernels -- every shipped config): EVERYTHING runs in librcf_loss.so as one
    autograd node (fused_head.RcfHeadFn): conv stem (mma.sync 3xTF32), second conv + its data and weight gradients on
    tcgen05 / TMA / tensor memory (csrc/rcf_conv64*.cu), pooling, segment MLP, fit, loss and the whole backward;
  * other head shapes: the second conv of ``flow_feat_before_agg`` (reference :89-91) goes through torch (cuDNN), the
    rest (mask normalisation, masked pooling, segment MLP, affine / quadratic fit, residual, reconstruction, loss, and
    the whole backward) through librcf_loss.so (sm_100a kernels).
There is no PyTorch implementation of the loss in this package: without the CUDA library (or on CPU
tensors) ``forward`` raises.

Documented deviations from the reference:
  * the returned visualisation ``flows`` are detached (the reference returns graph-attached tensors,
    but its only caller uses them for JPEG dumps, models/rcf_model.py:446-460);
  * ``coord_map`` is not materialised (the kernels generate coordinates on the fly), so no
    hard-coded ``.cuda()`` at construction (reference :143,:146);
  * no gradient is produced for the ground-truth flows.
"""
from __future__ import annotations

import logging
from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .function import LossSpec, rcf_motion_loss
from .fused_head import RcfHeadFn
from .conv64 import conv64_supported, default_nprod
from .resize import resize_bilinear_multi
from .stem import flow_stem, stem_supported

logger = logging.getLogger("main")


class Objectview(object):
    """dict -> attribute shim kept for import compatibility (reference :10-15)."""

    def __init__(self, d):
        self.__dict__ = d

    def keys(self):
        return self.__dict__.keys()


def get_norm_flow(lis1, lis2):
    """Visualisation scaling of two [B,2,H,W] flows (reference :18-30): ch0 / (H/2), ch1 / (W/2)."""
    def norm(x):
        _h, _w = x.shape[2:]
        return _h, _w, torch.cat([x[:, 0:1] / (_h / 2.0), x[:, 1:2] / (_w / 2.0)], 1)

    _, _, flow = norm(lis1)
    _h, _w, flow2 = norm(lis2)
    return _h, _w, flow, flow2


class FlowAggregationHeadWithResidual(nn.Module):
    """RCF motion-loss head; see module docstring.  Constructor mirrors reference :50-74."""

    def __init__(self,
                 args,
                 ssim_sz=1,
                 mask_layer=5,
                 create_flownet=False,
                 flow_feat_before_agg_kernel_size=3,
                 num_flow_feat_channels=64,
                 outlier_robust_loss=False,
                 eps=0.01,
                 q=0.4,
                 mask_size=(48, 48),
                 residual_adjustment_scale=10.,
                 norm_flow=False,
                 clamp_flow_t=None,
                 filter_flow_t=None,
                 free_residual=False,
                 free_residual_with_affine=False,
                 free_residual_with_affine_quadratic=False,
                 object_free_residual=False,
                 free_scale=False,
                 affine_residual=False,
                 allow_residual_resize=False,
                 pred_div_coeff=10.):
        self.mask_layer = mask_layer
        super().__init__()
        logger.info("[info] ssim_sz={}".format(ssim_sz))
        self.args = args
        assert create_flownet            # reference :82

        ks = flow_feat_before_agg_kernel_size
        self.flow_feat_before_agg = nn.Sequential(
            nn.Conv2d(2, num_flow_feat_channels, kernel_size=ks, stride=1, dilation=1, padding=(ks - 1) // 2, bias=True),
            nn.LeakyReLU(0.1, inplace=True),
            nn.Conv2d(num_flow_feat_channels, num_flow_feat_channels, kernel_size=ks, stride=1, dilation=1,
                      padding=(ks - 1) // 2, bias=True),
            nn.LeakyReLU(0.1, inplace=True),
        )
        # Parameters of the segment MLP; applied inside the CUDA segment kernel (never called as modules).
        self.flow_feat_after_agg = nn.Sequential(
            nn.Conv1d(num_flow_feat_channels, num_flow_feat_channels, kernel_size=1, stride=1, dilation=1, bias=True),
            nn.LeakyReLU(0.1, inplace=True),
            nn.Conv1d(num_flow_feat_channels, 2, kernel_size=1, stride=1, dilation=1, bias=True),
        )
        self.num_flow_feat_channels = num_flow_feat_channels

        self.outlier_robust_loss = outlier_robust_loss
        if self.outlier_robust_loss:
            logger.info("Using outlier robust loss")
        self.eps = eps
        self.q = q
        self.mask_size = tuple(mask_size)
        self.residual_adjustment_scale = residual_adjustment_scale
        self.pred_div_coeff = pred_div_coeff
        logger.info(f"Prediction division coefficient: {self.pred_div_coeff}")
        self.norm_flow = norm_flow
        self.clamp_flow_t = clamp_flow_t
        self.filter_flow_t = filter_flow_t

        self.free_residual = free_residual
        self.free_residual_with_affine = free_residual_with_affine
        self.free_residual_with_affine_quadratic = free_residual_with_affine_quadratic
        if self.free_residual_with_affine_quadratic:
            assert self.free_residual_with_affine, \
                "free_residual_with_affine needs to be enabled to enable free_residual_with_affine_quadratic"
        self.object_free_residual = object_free_residual
        self.free_scale = free_scale
        self.affine_residual = affine_residual
        assert (int(self.free_residual) + int(self.free_residual_with_affine) + int(self.object_free_residual)
                + int(self.free_scale) + int(self.affine_residual)) <= 1, \
            f"Only one of {self.free_residual}, {self.free_residual_with_affine}, {self.object_free_residual}, " \
            f"{self.free_scale}, {self.affine_residual}"
        self.allow_residual_resize = allow_residual_resize
        # Extensions (not constructor arguments, so YAML configs of the reference keep working):
        #   return_flows=False skips the 4-5 visualisation tensors (the reference always builds them, :370-395);
        #   loss_inv_n > 0 replaces the mean's 1/(B*2*H*W) -- set it to 1/(B_global*2*H*W) when the batch is sharded
        #   by hand and the per-rank losses are meant to be SUMMED (distributed.global_inv_n); under DDP leave it 0.
        #   channels_last_features=False keeps the conv branch in NCHW (default: channels-last, see _features_preact).
        self.return_flows = True
        self.loss_inv_n = 0.0
        self.channels_last_features = True
        #   handwritten_stem=False sends the first conv + LeakyReLU through cuDNN/ATen instead of csrc/rcf_stem.cu.
        self.handwritten_stem = True
        #   tensor_core_conv=False sends the second conv through cuDNN/ATen instead of csrc/rcf_conv64*.cu (tcgen05).
        #   conv_precision: precision level of the tcgen05 convs; None = what the reference's own convs would do under the
        #   caller's torch settings: 3 (fp32-grade: three bf16 products per fp32 product) when
        #   torch.backends.cudnn.allow_tf32 is False and no autocast is active, otherwise 2 (TF32-class: one product of fp16
        #   operands, scaled fp16 gradient).  1 = plain bf16 operands (same cost as 2, 8-bit significands).
        self.tensor_core_conv = True
        self.conv_precision = None

    # ------------------------------------------------------------------------------------------
    @property
    def _D(self) -> int:
        if not self.free_residual_with_affine:
            return 0
        return 5 if self.free_residual_with_affine_quadratic else 2

    def _fused_clamp(self) -> bool:
        """Clamp can be fused into the kernels' loads unless norm/filter (torch path) are in play."""
        return not self.norm_flow and self.filter_flow_t is None

    def norm_and_clamp_flow(self, flow):
        """reference :150-162, including its in-place filter on the caller's tensor when no copy preceded."""
        if self.norm_flow:
            flow = flow / flow.abs().max()
        if self.clamp_flow_t is not None:
            flow = flow.clamp(min=-self.clamp_flow_t, max=self.clamp_flow_t)
        if self.filter_flow_t is not None:
            flow[flow.abs() < self.filter_flow_t] = 0.
        return flow

    def _check_residual_mode(self):
        if not (self.free_residual or self.free_residual_with_affine):
            # the reference reaches `return ..., residual_adjustment, ...` with the name unbound (:305-310)
            raise UnboundLocalError("local variable 'residual_adjustment' referenced before assignment")

    def _features_preact(self, flow, stem_flows=None, stem_clamp=None):
        """conv -> LeakyReLU -> conv of flow_feat_before_agg (reference :84-91); the trailing LeakyReLU (:92) is applied
        inside the pooling kernels (and its derivative inside the pooling backward), which saves one full read+write
        of the [B,Cf,H,W] map in forward and one read + one read+write in backward.

        `flow` is the (prepared) conv input [N,2,H,W]; when `stem_flows` is given (per-direction raw flows whose only
        preparation is the clamp `stem_clamp`) and the shape is supported, the first conv + LeakyReLU run as the
        hand-written stem kernel instead and `flow` may be None."""
        seq = self.flow_feat_before_agg
        if self._use_channels_last() and stem_flows is not None and self.handwritten_stem \
                and stem_supported(self.num_flow_feat_channels, seq[0].kernel_size[0]) and seq[0].bias is not None:
            c2 = seq[2]
            act1 = flow_stem(stem_flows, seq[0].weight, seq[0].bias, stem_clamp, seq[1].negative_slope)
            feat = self._conv2(act1, c2)
            if c2.bias is None:
                return feat, None
            if feat.is_contiguous(memory_format=torch.channels_last):
                return feat, c2.bias
            return feat + c2.bias.view(1, -1, 1, 1), None
        if flow is None:
            flow = torch.cat([self._clamped(f, stem_clamp) for f in stem_flows], 0) if len(stem_flows) > 1 \
                else self._clamped(stem_flows[0], stem_clamp)
        if self._use_channels_last():
            # cuDNN's tensor-core convolutions are channels-last natively: feeding them channels-last tensors removes
            # their NCHW<->NHWC transposes (27 of ~70 launches per step at 96x96); the pooling kernels read that layout.
            # The last conv runs WITHOUT its bias: the pooling kernels add it on load and return its gradient, which
            # removes ATen's separate bias-add and bias-gradient reduction kernels over the [B,Cf,H,W] map.
            flow = flow.contiguous(memory_format=torch.channels_last)
            c2 = seq[2]
            feat = self._conv2(seq[1](seq[0](flow)), c2)
            if c2.bias is None:
                return feat, None
            if feat.is_contiguous(memory_format=torch.channels_last):
                return feat, c2.bias
            return feat + c2.bias.view(1, -1, 1, 1), None       # cuDNN answered in NCHW: plain bias add
        return seq[2](seq[1](seq[0](flow))), None

    @staticmethod
    def _conv2(x, c2):
        """bias-free second conv (general head shapes: ATen / cuDNN)."""
        return F.conv2d(x, c2.weight, None, c2.stride, c2.padding, c2.dilation, c2.groups)

    @staticmethod
    def _clamped(f, t):
        return f if t is None else f.clamp(min=-t, max=t)

    def _use_channels_last(self) -> bool:
        Cf = self.num_flow_feat_channels
        return bool(self.channels_last_features) and Cf % 4 == 0 and Cf <= 128 and 256 % (Cf // 4) == 0

    def _spec(self, K, H, W, *, want_vis, vis_norm, inv_n=0.0, clamp_fused=True) -> LossSpec:
        unbounded = bool(self.free_residual and self.residual_adjustment_scale == -1.)
        return LossSpec(K=K, H=H, W=W, D=self._D, Cf=self.num_flow_feat_channels,
                        robust=bool(self.outlier_robust_loss), eps=float(self.eps), q=float(self.q),
                        resid_scale=float(self.residual_adjustment_scale), pred_div=float(self.pred_div_coeff),
                        clamp_t=(self.clamp_flow_t if clamp_fused else None), unbounded_residual=unbounded,
                        inv_n=inv_n, want_vis=want_vis,
                        vis_scale=((2.0 / H, 2.0 / W) if vis_norm else (1.0, 1.0)),
                        feat_lrelu_slope=float(self.flow_feat_before_agg[3].negative_slope))

    def _prepare(self, flow, resid, H, W):
        """Returns (flow for the kernels, flow for the conv branch, clamp_fused, residual at mask_size)."""
        if self._fused_clamp():
            k_flow, c_flow = flow, None        # the clamp is applied inside the kernels (loss passes and conv stem)
            fused = True
        else:
            k_flow = c_flow = self.norm_and_clamp_flow(flow)
            fused = False
        return k_flow, c_flow, fused, resid

    def _resize_residuals(self, resids):
        """:271-273, :294-296 -- F.interpolate(resid, mask_size, mode='bilinear'); both directions in one launch."""
        if not self.allow_residual_resize:
            return list(resids)
        todo = [i for i, r in enumerate(resids) if tuple(r.shape[-2:]) != self.mask_size]
        out = list(resids)
        if len(todo) == 2 and resids[0].shape == resids[1].shape:
            out[todo[0]], out[todo[1]] = resize_bilinear_multi([resids[0], resids[1]], self.mask_size)
        else:
            for i in todo:
                out[i] = resize_bilinear_multi([resids[i]], self.mask_size)[0]
        return out

    def _mlp_params(self):
        l0, l2 = self.flow_feat_after_agg[0], self.flow_feat_after_agg[2]
        return l0.weight, l0.bias, l2.weight, l2.bias

    def _tc_head_supported(self) -> bool:
        """Default head shape: the whole head runs as one autograd node on the library's own kernels (fused_head.py)."""
        seq = self.flow_feat_before_agg
        c1, c2 = seq[0], seq[2]
        return bool(self.tensor_core_conv and self.handwritten_stem and self.channels_last_features
                    and self.num_flow_feat_channels == 64 and self._fused_clamp()
                    and c1.bias is not None and c2.bias is not None and conv64_supported(c2)
                    and c1.kernel_size[0] in (1, 3, 5) and c1.kernel_size[0] == c1.kernel_size[1]
                    and tuple(c1.padding) == ((c1.kernel_size[0] - 1) // 2,) * 2 and tuple(c1.stride) == (1, 1)
                    and 0.0 <= seq[1].negative_slope <= 1.0)

    def _run(self, masks5, flows, resids, *, want_vis, vis_norm, inv_n=0.0):
        """masks5 [B,ndir,K,H,W]; flows / resids: per-direction lists."""
        self._check_residual_mode()
        if not masks5.is_cuda:
            raise RuntimeError("FlowAggregationHeadWithResidual (B200) needs CUDA tensors; there is no CPU fallback")
        B, ndir, K, H, W = masks5.shape
        in_autocast = torch.is_autocast_enabled()
        if self._tc_head_supported():
            with torch.autocast(device_type="cuda", enabled=False):
                resids = [r.float() for r in resids]
                same_lowres = (self.allow_residual_resize and tuple(resids[0].shape[-2:]) != (H, W)
                               and all(r.shape == resids[0].shape for r in resids))
                if not same_lowres:          # (all directions at one lower resolution: up-sampled inside the library call)
                    resids = self._resize_residuals(resids)
                    for r in resids:
                        assert r.shape[-2:] == (H, W), f"residual spatial size {tuple(r.shape[-2:])} != mask size {(H, W)}"
                spec = self._spec(K, H, W, want_vis=want_vis, vis_norm=vis_norm, inv_n=inv_n, clamp_fused=True)
                nprod = self.conv_precision if self.conv_precision is not None else default_nprod(in_autocast)
                seq = self.flow_feat_before_agg
                out = RcfHeadFn.apply(spec, int(nprod), float(seq[1].negative_slope), masks5.float(), seq[0].weight, seq[0].bias,
                                      seq[2].weight, seq[2].bias, *self._mlp_params(), *[f.float() for f in flows], *resids)
            return out[0], out[1], tuple(out[2:])
        with torch.autocast(device_type="cuda", enabled=False):
            masks5 = masks5.float()
            k_flows, c_flows, rs = [], [], []
            fused = True
            resids = self._resize_residuals([r.float() for r in resids])
            for flow, resid in zip(flows, resids):
                kf, cf_, fused_i, r = self._prepare(flow.float(), resid.float(), H, W)
                fused = fused and fused_i
                assert r.shape[-2:] == (H, W), f"residual spatial size {tuple(r.shape[-2:])} != mask size {(H, W)}"
                k_flows.append(kf.detach()); c_flows.append(cf_); rs.append(r)
            # one pass of the conv branch over both directions (batch-concatenated), then a free 5-D view
            if fused:   # only the clamp stands between the raw flow and the conv: the stem kernel applies it itself
                feat, feat_bias = self._features_preact(None, stem_flows=k_flows, stem_clamp=self.clamp_flow_t)
            else:
                feat, feat_bias = self._features_preact(torch.cat(c_flows, 0) if ndir > 1 else c_flows[0])
            assert feat.shape[2:] == masks5.shape[3:], \
                f"{feat.shape[2:]} != {masks5.shape[3:]} (should match on spatial dimension)"   # :247-248
            feat = feat.view(ndir, B, *feat.shape[1:])
            spec = self._spec(K, H, W, want_vis=want_vis, vis_norm=vis_norm, inv_n=inv_n, clamp_fused=fused)
            loss, total, vis = rcf_motion_loss(spec, masks5, k_flows, rs, feats=feat, mlp=self._mlp_params(),
                                               feat_bias=feat_bias, with_total=True)
        return loss, total, vis

    # ------------------------------------------------------------------------------------------
    def get_demean_affine_flow(self, mask, flow):
        """reference :164-233: the de-meaned affine / quadratic part of the fit, [B,2,H,W] (no grad).
        `flow` is used as given (the reference passes the already clamped flow)."""
        assert self.free_residual_with_affine
        B, K, H, W = mask.shape
        zeros = mask.new_zeros(B, 2 * K, H, W)
        theta = mask.new_zeros(B, 2, K)
        spec = LossSpec(K=K, H=H, W=W, D=self._D, Cf=0, clamp_t=None, want_vis=True)
        with torch.no_grad():
            _, vis = rcf_motion_loss(spec, mask.detach().unsqueeze(1), [flow.detach()], [zeros], thetas=[theta])
        return vis[4]

    def aggregate_flow_with_residual(self, mask, flow, all_pred_residual):
        """reference :235-310 -> (flow_overall, flow_agg, residual_adjustment, flow_affine), detached.
        `flow` is taken as already prepared (clamped), as in the reference's call sites (:344-347)."""
        self._check_residual_mode()
        B, K, H, W = mask.shape
        flow = flow.float()
        resid = self._resize_residuals([all_pred_residual.float()])[0]
        with torch.no_grad(), torch.autocast(device_type="cuda", enabled=False):
            feat, feat_bias = self._features_preact(None, stem_flows=[flow], stem_clamp=None)
            assert feat.shape[2:] == mask.shape[2:], f"{feat.shape[2:]} != {mask.shape[2:]} (should match on spatial dimension)"
            spec = self._spec(K, H, W, want_vis=True, vis_norm=False, clamp_fused=False)
            _, vis = rcf_motion_loss(spec, mask.float().unsqueeze(1), [flow], [resid], feats=feat.unsqueeze(0),
                                     mlp=self._mlp_params(), feat_bias=feat_bias)
        return vis[1], vis[2], vis[3], (vis[4] if len(vis) > 4 else None)

    def forward(self, imgs, masks, gt_fw_flows, gt_bw_flows, all_pred_residual_fw, all_pred_residual_bw):
        """reference :312-399.  masks [B,2,K,H,W]; gt_*_flows [B,1,2,H,W]; residuals [B,2K,h,w]."""
        return self._forward_impl(imgs, masks, gt_fw_flows, gt_bw_flows, all_pred_residual_fw, all_pred_residual_bw,
                                  want_flows=self.return_flows)

    def _forward_impl(self, imgs, masks, gt_fw_flows, gt_bw_flows, all_pred_residual_fw, all_pred_residual_bw, *, want_flows):
        """forward() with the visualisation switch as an argument (graphed.py captures with want_flows=False without
        touching the module's `return_flows` attribute, which other callers of the same head may be reading)."""
        flow_loss = {'seg_fw': 0., 'seg_bw': 0.}
        flows: Dict[str, List[torch.Tensor]] = {'gt_flow': [], 'pred_flow': [], 'agg_flow': [],
                                                'residual_adj': [], 'affine_flow': []}
        batch_size, im_num, _, im_h, im_w = imgs.shape
        assert im_num == 2, "Other im_num not implemented"           # :324

        gt_fw_flow = gt_fw_flows[:, 0, ...]
        gt_bw_flow = gt_bw_flows[:, 0, ...]
        loss, total, vis = self._run(masks, [gt_fw_flow, gt_bw_flow], [all_pred_residual_fw, all_pred_residual_bw],
                                     want_vis=want_flows, vis_norm=True, inv_n=float(self.loss_inv_n))
        flow_loss['seg_fw'] = loss[0]
        flow_loss['seg_bw'] = loss[1]
        if want_flows:
            flows['gt_flow'].append(vis[0])
            flows['pred_flow'].append(vis[1])
            flows['agg_flow'].append(vis[2])
            flows['residual_adj'].append(vis[3])
            if len(vis) > 4:
                flows['affine_flow'].append(vis[4])
        flow_loss['seg'] = total      # = seg_fw + seg_bw (:397), summed in fp32 by the library
        return flows, flow_loss
