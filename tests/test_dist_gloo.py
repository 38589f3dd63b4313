"""world_size-2 gloo test of the batch-sharding logic used by the multi-GPU path (bench.py --gpus N):
per-rank loss partials with the global normaliser all-reduce (SUM) to the global loss, and the
concatenated per-shard gradients equal the global gradients.  The per-rank compute here is the numpy
oracle (CPU stand-in for the kernels, tests only); the same identity is checked on the real kernels
in tests/test_gpu_parity.py::test_full_size_properties."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rcf_oracle as O
from rcf_unsupvideoseg_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, K, H, W = 5, 3, 6, 8
        cfg = O.OracleConfig(mask_layer=K, mask_size=(H, W), num_flow_feat_channels=4, free_residual_with_affine=True,
                             clamp_flow_t=20.0)
        params = O.init_params(cfg, seed=0)
        masks, fw, bw, rfw, rbw = [a.astype(np.float64) for a in O.synthetic_inputs(B, K, H, W, seed=1)]
        lo, hi = D.shard_bounds(B, rank, world)
        _, loss, caches = O.head_forward(masks[lo:hi], fw[lo:hi], bw[lo:hi], rfw[lo:hi], rbw[lo:hi], params, cfg)
        # the kernels take inv_n; the oracle normalises by the local count, so rescale to the global normaliser
        scale = (hi - lo) * 2 * H * W * D.global_inv_n(B, H, W)
        part = torch.tensor([loss["seg_fw"] * scale, loss["seg_bw"] * scale], dtype=torch.float64)
        work = D.all_reduce_loss(part, async_op=True)
        grads = O.head_backward(caches, params, gbar=scale)
        work.wait()
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, grads["d_masks"]))
        if rank == 0:
            _, loss_g, caches_g = O.head_forward(masks, fw, bw, rfw, rbw, params, cfg)
            g_g = O.head_backward(caches_g, params, gbar=1.0)
            dm = np.concatenate([g for _, _, g in sorted(gathered, key=lambda t: t[0])], axis=0)
            out["loss_err"] = abs(float(part.sum()) - loss_g["seg"]) / loss_g["seg"]
            out["grad_err"] = float(np.linalg.norm(dm - g_g["d_masks"]) / np.linalg.norm(g_g["d_masks"]))
    finally:
        dist.destroy_process_group()


def test_sharded_loss_and_grads_match_global():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert out["loss_err"] < 1e-12
        assert out["grad_err"] < 1e-12


def test_shard_bounds_cover_batch():
    for B in (1, 2, 7, 16, 64):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(4, 4, 4)
    assert D.all_reduce_loss(torch.zeros(2)) is None     # not initialised -> no-op
