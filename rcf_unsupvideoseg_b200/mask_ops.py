"""Caller-side mask preparation and mask losses of the reference's RCFModel.forward_train as one fused op each way.

    masks, losses = mask_losses(logits, compact_channel=0, pl_masks=pl, object_channel=oc, pl_pos_weight=.., ...)
    masks, entropy = softmax_entropy(logits)          # logits [B, I, K, H, W]

replaces (models/rcf_model.py, models/compactness_head.py)
    :433      all_pred_mask = F.softmax(all_pred_mask, dim=2)
    :434      log_all_pred_mask = F.log_softmax(all_pred_mask, dim=2)            # (sic) log-softmax of the probabilities
    :376-378  get_entropy_loss = -(all_pred_mask * log_all_pred_mask).sum(dim=2).mean()
    :350-374  get_sharpen_loss: KL to the sharpened masks (utils.sharpen) or the object-aware hinge
    :380-408  get_pl_loss / get_crf_loss: pos/neg weighted MSE of the object channel towards a (thresholded) target mask
    compactness_head.py:33-56  CompactnessHead.get_compactness_loss
`masks` goes to the motion loss (FlowAggregationHeadWithResidual); the losses are weighted and summed by the caller
(:470-500).  The backward merges the mask gradient coming back from the motion loss with the gradients of the mask losses
and applies the softmax backward in a single pass (csrc/rcf_maskops.cu).  CUDA only; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib


class _MaskLossesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, compact_channel: int, pl_channel: int, pl_th: float, wpos: float, wneg: float,
                sharpen_mode: int, sharpen_channel: int, t_sharpen: float):
        lib = _lib.load_library()
        if not logits.is_cuda:
            raise RuntimeError("mask_losses: CUDA tensors required (no CPU fallback)")
        assert logits.dim() == 5, "logits [B, I, K, H, W]"
        B, I, K, H, W = logits.shape
        x = logits.detach().float().contiguous()
        tgt = None
        if pl_channel >= 0:
            assert target is not None and tuple(target.shape) == (B, I, H, W), "target [B, I, H, W]"
            tgt = target.detach().float().contiguous()
        cfg = _lib.RcfMaskCfg(B * I, K, H, W, compact_channel, pl_channel, pl_th, wpos, wneg, sharpen_mode, sharpen_channel,
                              t_sharpen)
        masks = torch.empty_like(x)
        losses = torch.empty(4, dtype=torch.float32, device=x.device)
        fstats = torch.empty(B * I, 2, dtype=torch.float32, device=x.device)
        n = C.c_size_t()
        _lib.check(lib.rcf_mask_prep_workspace_floats(B * I, H * W, C.byref(n)), "rcf_mask_prep_workspace_floats")
        ws = torch.empty(n.value, dtype=torch.float32, device=x.device)
        with _lib.device_guard(x.device):
            _lib.check(lib.rcf_mask_losses_forward(C.byref(cfg), x.data_ptr(), tgt.data_ptr() if tgt is not None else None,
                                                   masks.data_ptr(), losses.data_ptr(), fstats.data_ptr(), ws.data_ptr(),
                                                   _lib.raw_stream(x.device)), "rcf_mask_losses_forward")
        ctx.cfg = cfg
        ctx.save_for_backward(masks, fstats, *([tgt] if tgt is not None else []))
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(fstats)
        return masks, losses, fstats

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_masks, g_losses, _g_fstats):
        lib = _lib.load_library()
        masks, fstats, *rest = ctx.saved_tensors
        if g_masks is None and g_losses is None:
            return (None,) * 10
        gm = g_masks.float().contiguous() if g_masks is not None else None
        gl = g_losses.detach().float().contiguous() if g_losses is not None else None
        dl = torch.empty_like(masks)
        with _lib.device_guard(masks.device):
            _lib.check(lib.rcf_mask_losses_backward(C.byref(ctx.cfg), masks.data_ptr(), rest[0].data_ptr() if rest else None,
                                                    gm.data_ptr() if gm is not None else None,
                                                    gl.data_ptr() if gl is not None else None, fstats.data_ptr(),
                                                    dl.data_ptr(), _lib.raw_stream(masks.device)),
                       "rcf_mask_losses_backward")
        return (dl,) + (None,) * 9


def mask_losses(logits: torch.Tensor, *, compact_channel: Optional[int] = None, pl_masks: Optional[torch.Tensor] = None,
                object_channel: Optional[int] = None, pl_mask_pos_th: float = -1.0, pl_pos_weight: float = 1.0,
                pl_neg_weight: float = 1.0, sharpen: Optional[str] = None, t_sharpen: float = 0.25
                ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """logits [B, I, K, H, W] -> (masks, {'entropy', 'compactness', 'pl', 'sharpen'}).

    compact_channel: channel of CompactnessHead (None: loss not computed, entry is 0; -1: the object channel, as in
    models/compactness_head.py:19-27 -- object_channel must then be given).  Channel indices outside [0, K) raise.
    pl_masks [B, I, H, W] + object_channel: PL (or CRF) target and the object channel; pl_mask_pos_th = -1 uses the target
    as it is, any other value binarises it (`pl_masks > th`), exactly as get_pl_loss / get_crf_loss do.
    sharpen: None | 'kl' (get_sharpen_loss with object_aware_sharpening=False: KL to utils.sharpen(masks, t_sharpen)) |
    'object_hinge' (object_aware_sharpening=True; needs object_channel).
    """
    assert sharpen in (None, 'kl', 'object_hinge')
    smode = {None: 0, 'kl': 1, 'object_hinge': 2}[sharpen]
    if smode == 2:
        assert object_channel is not None, "sharpen='object_hinge' needs object_channel"
    if compact_channel is not None and int(compact_channel) == -1:
        if object_channel is None:
            raise ValueError("compact_channel=-1 means 'the object channel' (compactness_head.py:19-27): pass object_channel")
        compact_channel = object_channel
    K = logits.shape[2]
    for name, ch in (("compact_channel", compact_channel), ("object_channel", object_channel)):
        if ch is not None and not 0 <= int(ch) < K:
            raise ValueError(f"{name}={ch} is outside [0, {K})")
    use_pl = pl_masks is not None and object_channel is not None
    masks, losses, _ = _MaskLossesFn.apply(logits, pl_masks if use_pl else None,
                                           -1 if compact_channel is None else int(compact_channel),
                                           int(object_channel) if use_pl else -1, float(pl_mask_pos_th),
                                           float(pl_pos_weight), float(pl_neg_weight), smode,
                                           int(object_channel) if smode == 2 else -1, float(t_sharpen))
    return masks, {"entropy": losses[0], "compactness": losses[1], "pl": losses[2], "sharpen": losses[3]}


def softmax_entropy(logits: torch.Tensor):
    """logits [B, I, K, H, W] -> (masks = softmax over K, entropy loss of models/rcf_model.py:376-378)."""
    masks, d = mask_losses(logits)
    return masks, d["entropy"]
