"""Import-compatible mirror of the reference's utils/loss_utils.py (:7-108): mmseg-style loss reduction helpers
and the PAWS ``sharpen``.  Pure tensor glue (a handful of elementwise ops on already-reduced or mask-sized
tensors); the RCF head itself never calls them -- ``sharpen`` feeds the caller's
KL-sharpening loss (models/rcf_model.py:371-373)."""
from __future__ import annotations

import functools

import torch.nn.functional as F


def reduce_loss(loss, reduction):
    """'none' | 'mean' | 'sum'  (reference :7-24)."""
    mode = F._Reduction.get_enum(reduction)
    if mode == 0:
        return loss
    return loss.mean() if mode == 1 else loss.sum()


def weight_reduce_loss(loss, weight=None, reduction='mean', avg_factor=None):
    """Element-wise weight, then reduce; `avg_factor` replaces the mean's denominator  (reference :27-56)."""
    if weight is not None:
        assert weight.dim() == loss.dim()
        if weight.dim() > 1:
            assert weight.size(1) == 1 or weight.size(1) == loss.size(1)
        loss = loss * weight
    if avg_factor is None:
        return reduce_loss(loss, reduction)
    if reduction == 'mean':
        return loss.sum() / avg_factor
    if reduction != 'none':
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return loss


def weighted_loss(loss_func):
    """Decorator adding (weight, reduction, avg_factor) to an element-wise loss  (reference :59-102)."""
    @functools.wraps(loss_func)
    def wrapper(pred, target, weight=None, reduction='mean', avg_factor=None, **kwargs):
        return weight_reduce_loss(loss_func(pred, target, **kwargs), weight, reduction, avg_factor)
    return wrapper


def sharpen(p, T, dim=1):
    """p ** (1/T), renormalised along `dim`  (reference :105-108)."""
    q = p ** (1. / T)
    return q / q.sum(dim=dim, keepdim=True)
