"""Caller-side mask preparation of the reference's RCFModel.forward_train as one fused op each way.

    masks, entropy = softmax_entropy(logits)          # logits [B, I, K, H, W]

replaces (models/rcf_model.py)
    :433      all_pred_mask = F.softmax(all_pred_mask, dim=2)
    :434      log_all_pred_mask = F.log_softmax(all_pred_mask, dim=2)            # (sic) log-softmax of the probabilities
    :376-378  get_entropy_loss = -(all_pred_mask * log_all_pred_mask).sum(dim=2).mean()
`masks` goes to the motion loss (FlowAggregationHeadWithResidual), `entropy * w_entropy` is added to the total loss
(:476-478; configs/rcf/rcf_stage1.yaml:67).  The backward merges the mask gradient coming back from the motion loss and the
entropy gradient and applies the softmax backward in a single pass (csrc/rcf_maskops.cu).  CUDA only; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class _SoftmaxEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        lib = _lib.load_library()
        if not logits.is_cuda:
            raise RuntimeError("softmax_entropy: CUDA tensors required (no CPU fallback)")
        assert logits.dim() == 5, "logits [B, I, K, H, W]"
        B, I, K, H, W = logits.shape
        x = logits.detach().float().contiguous()
        masks = torch.empty_like(x)
        ent = torch.empty((), dtype=torch.float32, device=x.device)
        n = C.c_size_t()
        _lib.check(lib.rcf_mask_prep_workspace_floats(B * I, H * W, C.byref(n)), "rcf_mask_prep_workspace_floats")
        ws = torch.empty(n.value, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.rcf_mask_prep_forward(x.data_ptr(), masks.data_ptr(), ent.data_ptr(), ws.data_ptr(), B * I, K, H * W,
                                                 torch.cuda.current_stream(x.device).cuda_stream), "rcf_mask_prep_forward")
        ctx.save_for_backward(masks)
        ctx.set_materialize_grads(False)
        return masks, ent

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_masks, g_ent):
        lib = _lib.load_library()
        (masks,) = ctx.saved_tensors
        if g_masks is None and g_ent is None:
            return None
        B, I, K, H, W = masks.shape
        gm = g_masks.float().contiguous() if g_masks is not None else None
        ge = g_ent.detach().float().reshape(1).contiguous() if g_ent is not None else None
        dl = torch.empty_like(masks)
        with torch.cuda.device(masks.device):
            _lib.check(lib.rcf_mask_prep_backward(masks.data_ptr(), gm.data_ptr() if gm is not None else None,
                                                  ge.data_ptr() if ge is not None else None, dl.data_ptr(), B * I, K, H * W,
                                                  torch.cuda.current_stream(masks.device).cuda_stream), "rcf_mask_prep_backward")
        return dl


def softmax_entropy(logits: torch.Tensor):
    """logits [B, I, K, H, W] -> (masks = softmax over K, entropy loss of models/rcf_model.py:376-378)."""
    return _SoftmaxEntropyFn.apply(logits)
