"""CUDA-graph capture of the head's forward+backward for fixed shapes (the 48x48 / 96x96 training regime is
launch-bound: ~40 kernels of a few microseconds each; one graph launch replaces them).

    graphed = make_graphed_head(head, (imgs, masks, gt_fw, gt_bw, res_fw, res_bw))
    flow_loss = graphed(masks, gt_fw, gt_bw, res_fw, res_bw)        # {'seg_fw','seg_bw','seg'}; .backward() works

Uses torch.cuda.make_graphed_callables (streams + graphs, no tracing compiler).  The visualisation flows are
not produced on this path (they are only needed every `log_interval` iterations, models/rcf_model.py:456-460 --
call the plain head then).  As with any torch CUDA-graph capture, no autograd graph involving the head's
parameters may be alive while capturing (a live graph pins the parameters' AccumulateGrad nodes to the default
stream and the capture is invalidated): capture before training starts or after `loss.backward()` + `del loss`.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _LossOnly(nn.Module):
    def __init__(self, head, imgs_shape):
        super().__init__()
        self.head = head
        self._imgs = torch.empty(imgs_shape, device="meta")      # only .shape is read by forward (reference :322-324)

    def forward(self, masks, gt_fw, gt_bw, res_fw, res_bw):
        _, fl = self.head._forward_impl(self._imgs, masks, gt_fw, gt_bw, res_fw, res_bw, want_flows=False)
        return fl["seg_fw"], fl["seg_bw"], fl["seg"]


def make_graphed_head(head, sample_inputs, num_warmup_iters: int = 3):
    """sample_inputs = (imgs, masks, gt_fw_flows, gt_bw_flows, res_fw, res_bw) with the shapes/dtypes/devices and
    requires_grad flags of the real calls.  Returns a callable(masks, gt_fw, gt_bw, res_fw, res_bw) -> dict."""
    imgs, masks, gt_fw, gt_bw, res_fw, res_bw = sample_inputs
    mod = _LossOnly(head, tuple(imgs.shape))
    samples = tuple(t.detach().clone().requires_grad_(t.requires_grad) for t in (masks, gt_fw, gt_bw, res_fw, res_bw))
    graphed = torch.cuda.make_graphed_callables(mod, samples, num_warmup_iters=num_warmup_iters)

    def call(masks, gt_fw, gt_bw, res_fw, res_bw):
        seg_fw, seg_bw, seg = graphed(masks, gt_fw, gt_bw, res_fw, res_bw)
        return {"seg_fw": seg_fw, "seg_bw": seg_bw, "seg": seg}

    return call
