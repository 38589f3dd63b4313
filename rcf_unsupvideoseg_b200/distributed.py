"""Batch sharding of the RCF motion loss across GPUs (one process per GPU, NCCL over NVLink).

The loss shards naturally: every statistic is per (sample, direction, segment); the
only cross-sample coupling is the normaliser N = B_global * 2 * H * W of the mean (reference :361).
Each rank therefore runs the kernels on its contiguous slice of the batch with ``inv_n = 1 / N_global``;
the per-rank loss partials then SUM to the reference's global loss, and mask / residual gradients need no
communication at all.  The only collective is one all-reduce(SUM) of the 2-float loss vector (plus the
42 434-element parameter gradient when the head trains, which DDP already does: main.py:455).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of the batch dim; the first ``batch % world`` ranks get one extra sample."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def global_inv_n(batch_global: int, H: int, W: int) -> float:
    """1 / (elements in one direction's mean over the GLOBAL batch) -- LossSpec.inv_n for sharded runs."""
    return 1.0 / (float(batch_global) * 2.0 * H * W)


def all_reduce_loss(loss_partial: torch.Tensor, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
    """Sum the per-rank loss partials ([ndir] fp32) in place.  Returns the work handle when async."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(loss_partial, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
