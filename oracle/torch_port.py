"""CPU baseline port of the reference head in plain PyTorch ops (fp32, autograd backward).

TEST / BENCH INFRASTRUCTURE ONLY -- used by ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``)
and by tests as a second checker.  /root/reference cannot travel to the GPU box, so this module
restates the reference's op sequence (models/flow_aggregation_head_with_residual.py:150-399) with the
same tensor temporaries -- including the [B,Cf,K,H,W] broadcast product of :251-256 and the
[B,K,HW,2,D] outer products of :198-209 -- so that its CPU cost is the reference's CPU cost.
It is pinned to the reference's own fp32 outputs by tests/test_torch_port.py (golden fixtures).
Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class PortedHead(nn.Module):
    def __init__(self, mask_layer=5, flow_feat_before_agg_kernel_size=3, num_flow_feat_channels=64,
                 outlier_robust_loss=False, eps=0.01, q=0.4, mask_size=(48, 48), residual_adjustment_scale=10.,
                 norm_flow=False, clamp_flow_t=None, filter_flow_t=None, free_residual=False,
                 free_residual_with_affine=False, free_residual_with_affine_quadratic=False,
                 allow_residual_resize=False, pred_div_coeff=10.):
        super().__init__()
        c, k = num_flow_feat_channels, flow_feat_before_agg_kernel_size
        self.flow_feat_before_agg = nn.Sequential(
            nn.Conv2d(2, c, k, padding=(k - 1) // 2), nn.LeakyReLU(0.1, inplace=True),
            nn.Conv2d(c, c, k, padding=(k - 1) // 2), nn.LeakyReLU(0.1, inplace=True))
        self.flow_feat_after_agg = nn.Sequential(
            nn.Conv1d(c, c, 1), nn.LeakyReLU(0.1, inplace=True), nn.Conv1d(c, 2, 1))
        self.K = mask_layer
        self.robust, self.eps, self.q = outlier_robust_loss, eps, q
        self.mask_size = tuple(mask_size)
        self.scale, self.div = residual_adjustment_scale, pred_div_coeff
        self.norm_flow, self.clamp_t, self.filter_t = norm_flow, clamp_flow_t, filter_flow_t
        self.free, self.affine, self.quadratic = free_residual, free_residual_with_affine, free_residual_with_affine_quadratic
        self.resize = allow_residual_resize
        assert self.free or self.affine
        if self.affine:
            H, W = self.mask_size
            rr, cc = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
            cols = [rr, cc] + ([rr * rr, cc * cc, rr * cc] if self.quadratic else [])
            self.register_buffer("coords", torch.stack(cols, dim=2).reshape(H * W, -1).float(), persistent=False)

    def prepare(self, flow):                                            # :150-162
        if self.norm_flow:
            flow = flow / flow.abs().max()
        if self.clamp_t is not None:
            flow = flow.clamp(-self.clamp_t, self.clamp_t)
        if self.filter_t is not None:
            flow = torch.where(flow.abs() < self.filter_t, torch.zeros_like(flow), flow)
        return flow

    def affine_part(self, mask, flow):                                  # :164-233
        B, K, H, W = mask.shape
        w = (mask / mask.sum(dim=(2, 3), keepdim=True)).flatten(2, 3)            # [B,K,P]
        Fp = flow.flatten(2, 3).transpose(1, 2)                                   # [B,P,2]
        mF = torch.bmm(w, Fp)                                                     # [B,K,2]
        mu = w @ self.coords                                                      # [B,K,D]
        fd = Fp[:, None] - mF[:, :, None]                                         # [B,K,P,2]
        ud = self.coords[None, None] - mu[:, :, None]                             # [B,K,P,D]
        sFu = torch.einsum("bkp,bkpcd->bkcd", w, torch.einsum("bkpc,bkpd->bkpcd", fd, ud))
        suu = torch.einsum("bkp,bkpde->bkde", w, torch.einsum("bkpd,bkpe->bkpde", ud, ud))
        A = torch.linalg.solve(suu.float(), sFu.transpose(2, 3).float()).transpose(2, 3)   # [B,K,2,D]
        per_seg = torch.einsum("bkcd,bkpd->bkpc", A, ud).view(B, K, H, W, 2)
        return torch.einsum("bkhw,bkhwc->bchw", mask, per_seg)

    def direction(self, mask, flow, resid):                             # :235-310
        B, K, H, W = mask.shape
        wn = mask / mask.view(B, K, H * W, 1).sum(dim=2, keepdim=True)
        feat = self.flow_feat_before_agg(flow)
        pooled = (feat[:, :, None] * wn[:, None]).flatten(3, 4).sum(dim=-1)      # [B,Cf,K]  (:251-256)
        theta = self.flow_feat_after_agg(pooled)                                  # [B,2,K]
        agg = (theta[..., None, None] * mask[:, None]).sum(dim=2)
        if self.resize and tuple(resid.shape[-2:]) != self.mask_size:
            resid = F.interpolate(resid, self.mask_size, mode="bilinear")
        r = resid.unflatten(1, (2, K))
        if self.free and self.scale == -1.:
            res = (r * mask[:, None]).sum(dim=2)
        else:
            res = (torch.tanh(r / self.div) * mask[:, None]).sum(dim=2) * self.scale
        aff = self.affine_part(mask, flow) if self.affine else None
        pred = agg + res if aff is None else agg + aff + res
        return pred, agg, res, aff

    def forward(self, imgs, masks, gt_fw_flows, gt_bw_flows, res_fw, res_bw):  # :312-399
        assert imgs.shape[1] == 2
        out_flows = {"gt_flow": [], "pred_flow": [], "agg_flow": [], "residual_adj": [], "affine_flow": []}
        fw = self.prepare(gt_fw_flows[:, 0])
        bw = self.prepare(gt_bw_flows[:, 0])
        pf, af, rf, xf = self.direction(masks[:, 0], fw, res_fw)
        pb, ab, rb, xb = self.direction(masks[:, 1], bw, res_bw)

        def term(f, p):
            d = (f - p).abs().view(-1)
            return ((d + self.eps) ** self.q).mean() if self.robust else d.mean()

        loss = {"seg_fw": term(fw, pf), "seg_bw": term(bw, pb)}
        loss["seg"] = loss["seg_fw"] + loss["seg_bw"]

        def vis(a, b):
            H, W = a.shape[2:]
            s = a.new_tensor([2.0 / H, 2.0 / W]).view(1, 2, 1, 1)
            return torch.cat([a * s, b * s], dim=1)

        out_flows["gt_flow"].append(vis(fw, bw))
        out_flows["pred_flow"].append(vis(pf, pb))
        out_flows["agg_flow"].append(vis(af, ab))
        out_flows["residual_adj"].append(vis(rf, rb))
        if xf is not None:
            out_flows["affine_flow"].append(vis(xf, xb))
        return out_flows, loss


def synthetic_inputs(B, K, H, W, seed=0, device="cpu"):
    """SURVEY.md 8(d) inputs from torch's CPU generator (bit-reproducible on any box with this torch)."""
    g = torch.Generator().manual_seed(seed)
    masks = torch.softmax(torch.randn(B, 2, K, H, W, generator=g) * 2.0, dim=2)
    fw = torch.randn(B, 1, 2, H, W, generator=g) * 8.0
    bw = torch.randn(B, 1, 2, H, W, generator=g) * 8.0
    rfw = torch.randn(B, 2 * K, H, W, generator=g) * 5.0
    rbw = torch.randn(B, 2 * K, H, W, generator=g) * 5.0
    return [t.to(device) for t in (masks, fw, bw, rfw, rbw)]
