"""utils/warp_utils.py and utils/loss_utils.py mirrors: oracle pinned to the reference's outputs (CPU), host-side
helpers (CPU), and the CUDA sampling kernels against both (GPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_l2
from oracle import warp_oracle as WO


def _load(name):
    # the loss-path fixtures are collected by glob("*.npz") + their "meta" entry; these two files are separate
    z = np.load(os.path.join(GOLDEN_DIR, "aux", name))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def gw():
    return _load("warp_ops.npz")


def test_warp_oracle_matches_reference(gw):
    for pad in ("border", "zeros"):
        assert rel_l2(WO.flow_warp(gw["x"], gw["flow"], pad), gw[f"warp.{pad}"]) < 1e-12
        gx, gf = WO.flow_warp_backward(gw["x"], gw["flow"], gw["gout"], pad)
        assert rel_l2(gx, gw[f"warp.{pad}.gx"]) < 1e-12
        assert rel_l2(gf, gw[f"warp.{pad}.gflow"]) < 1e-12
    B, _, H, W = gw["flow21"].shape
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    grid = np.stack([xs, ys])[None] + gw["flow21"]
    assert rel_l2(WO.get_corresponding_map(grid), gw["corr_map"]) < 1e-12
    assert np.array_equal(WO.get_occu_mask_backward(gw["flow21"], 0.2), gw["occ_backward"])
    assert np.array_equal(WO.get_occu_mask_bidirection(gw["flow"], gw["flow21"]), gw["occ_bidirection"])


def test_host_helpers_match_reference(gw):
    from rcf_unsupvideoseg_b200 import loss_utils as LU
    from rcf_unsupvideoseg_b200 import warp_utils as WU
    assert np.array_equal(WU.mesh_grid(2, 3, 4).numpy(), gw["mesh_grid"])
    assert np.allclose(WU.norm_grid(WU.mesh_grid(2, 3, 4).double()).numpy(), gw["norm_grid"], atol=1e-15)
    lo = _load("loss_utils.npz")
    p = torch.from_numpy(lo["p"])
    assert np.allclose(LU.sharpen(p, 0.25, dim=1).numpy(), lo["sharpen_T0.25"], rtol=1e-13)
    loss, w = torch.from_numpy(lo["loss"]), torch.from_numpy(lo["weight"])
    assert np.allclose(LU.weight_reduce_loss(loss, w, "mean").numpy(), lo["wrl.mean"], rtol=1e-14)
    assert np.allclose(LU.weight_reduce_loss(loss, w, "sum").numpy(), lo["wrl.sum"], rtol=1e-14)
    assert np.allclose(LU.weight_reduce_loss(loss, w, "none").numpy(), lo["wrl.none"], rtol=1e-14)
    assert np.allclose(LU.weight_reduce_loss(loss, w, "mean", avg_factor=7.0).numpy(), lo["wrl.avg"], rtol=1e-14)
    with pytest.raises(ValueError):
        LU.weight_reduce_loss(loss, w, "sum", avg_factor=2.0)

    @LU.weighted_loss
    def l1(pred, target):
        return (pred - target).abs()
    pr, tg, ww = torch.tensor([0., 2., 3.]), torch.tensor([1., 1., 1.]), torch.tensor([1., 0., 1.])
    assert float(l1(pr, tg)) == pytest.approx(4 / 3) and float(l1(pr, tg, ww)) == pytest.approx(1.0)
    assert float(l1(pr, tg, ww, avg_factor=2)) == pytest.approx(1.5)
    assert LU.reduce_loss(pr, "none") is pr


def test_cpu_tensors_raise():
    from rcf_unsupvideoseg_b200 import warp_utils as WU
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        WU.flow_warp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(NotImplementedError):
        WU.flow_warp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4), mode="nearest")


@pytest.mark.gpu
def test_flow_warp_kernels_match_reference(gw):
    from rcf_unsupvideoseg_b200 import warp_utils as WU
    for pad in ("border", "zeros"):
        x = torch.from_numpy(gw["x"]).float().cuda().requires_grad_(True)
        fl = torch.from_numpy(gw["flow"]).float().cuda().requires_grad_(True)
        y = WU.flow_warp(x, fl, pad=pad)
        y.backward(torch.from_numpy(gw["gout"]).float().cuda())
        torch.cuda.synchronize()
        assert rel_l2(y.detach().cpu().numpy(), gw[f"warp.{pad}"]) < 2e-6
        assert rel_l2(x.grad.cpu().numpy(), gw[f"warp.{pad}.gx"]) < 2e-6
        assert rel_l2(fl.grad.cpu().numpy(), gw[f"warp.{pad}.gflow"]) < 2e-5


@pytest.mark.gpu
def test_corresponding_map_and_occlusion_masks(gw):
    from rcf_unsupvideoseg_b200 import warp_utils as WU
    f21 = torch.from_numpy(gw["flow21"]).float().cuda()
    B, _, H, W = f21.shape
    grid = WU.mesh_grid(B, H, W).cuda().float() + f21
    cm = WU.get_corresponding_map(grid)
    assert cm.shape == (B, 1, H, W)
    assert rel_l2(cm.cpu().numpy(), gw["corr_map"]) < 2e-6
    assert torch.equal(cm, WU.get_corresponding_map(grid))          # order-independent accumulation
    ob = WU.get_occu_mask_backward(f21, th=0.2).cpu().numpy()
    assert (ob != gw["occ_backward"]).mean() < 0.01                 # threshold ties only
    f12 = torch.from_numpy(gw["flow"]).float().cuda()
    obi = WU.get_occu_mask_bidirection(f12, f21).cpu().numpy()
    assert (obi != gw["occ_bidirection"]).mean() < 0.01


@pytest.mark.gpu
def test_flow_warp_large_random_vs_oracle():
    from rcf_unsupvideoseg_b200 import warp_utils as WU
    rng = np.random.default_rng(3)
    B, C, H, W = 2, 5, 60, 77
    x = rng.normal(size=(B, C, H, W)).astype(np.float32)
    fl = (rng.normal(size=(B, 2, H, W)) * 6).astype(np.float32)
    go = rng.normal(size=(B, C, H, W)).astype(np.float32)
    for pad in ("border", "zeros"):
        xt = torch.from_numpy(x).cuda().requires_grad_(True)
        ft = torch.from_numpy(fl).cuda().requires_grad_(True)
        y = WU.flow_warp(xt, ft, pad=pad)
        y.backward(torch.from_numpy(go).cuda())
        gx, gf = WO.flow_warp_backward(x, fl, go, pad)
        assert rel_l2(y.detach().cpu().numpy(), WO.flow_warp(x, fl, pad)) < 2e-6
        assert rel_l2(xt.grad.cpu().numpy(), gx) < 2e-6
        assert rel_l2(ft.grad.cpu().numpy(), gf) < 2e-5
