"""Caller-side mask preparation and mask losses: oracle vs fixtures generated from the unmodified
reference functions (tests/golden/make_golden_maskops.py) on CPU, kernels vs the oracle and the fixtures on the GPU."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN_DIR, rel_l2
from oracle import rcf_oracle as O

SHAPES = [(2, 2, 4, 12, 16), (1, 2, 3, 7, 9), (2, 2, 8, 6, 10), (1, 1, 1, 4, 4), (3, 2, 5, 5, 5)]
CASES = ("k4", "k3_th", "k5")


@pytest.fixture(scope="module")
def gm():
    z = np.load(os.path.join(GOLDEN_DIR, "aux", "mask_losses.npz"))
    return {k: z[k] for k in z.files}


def _cfg(gm, n):
    compact, oc, th, wp, wn, c0, c1, c2 = gm[f"{n}.cfg"]
    return int(compact), int(oc), float(th), float(wp), float(wn), (float(c0), float(c1), float(c2))


@pytest.mark.parametrize("n", CASES)
def test_mask_losses_oracle_matches_reference(gm, n):
    compact, oc, th, wp, wn, coef = _cfg(gm, n)
    m, losses, _ = O.mask_losses_forward(gm[f"{n}.logits"], compact, gm[f"{n}.pl"], oc, th, wp, wn)
    assert rel_l2(m, gm[f"{n}.masks"]) < 1e-13
    ref = gm[f"{n}.losses"]
    for i, k in enumerate(("entropy", "compactness", "pl")):
        assert abs(losses[k] - ref[i]) <= 1e-12 * abs(ref[i]) + 1e-15, k
    g = O.mask_losses_backward(m, gm[f"{n}.w_mask"], coef[0], coef[1], coef[2], compact, gm[f"{n}.pl"], oc, th, wp, wn)
    assert rel_l2(g, gm[f"{n}.dlogits"]) < 1e-11


@pytest.mark.parametrize("n", CASES)
@pytest.mark.parametrize("mode", ["kl", "object_hinge"])
def test_sharpen_oracle_matches_reference(gm, n, mode):
    """RCFModel.get_sharpen_loss (:350-374), both live variants, from the unmodified reference function"""
    compact, oc, th, wp, wn, coef = _cfg(gm, n)
    m, losses, _ = O.mask_losses_forward(gm[f"{n}.logits"], object_channel=oc, sharpen=mode, t_sharpen=0.25)
    ref = float(gm[f"{n}.sharpen.{mode}"])
    assert abs(losses["sharpen"] - ref) <= 1e-12 * abs(ref) + 1e-15
    g = O.mask_losses_backward(m, object_channel=oc, g_sharpen=1.7, sharpen=mode, t_sharpen=0.25)
    assert rel_l2(g, gm[f"{n}.sharpen.{mode}.dlogits"]) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("n", CASES)
@pytest.mark.parametrize("mode", ["kl", "object_hinge"])
def test_sharpen_kernels_match_reference_fixtures(gm, n, mode):
    from rcf_unsupvideoseg_b200.mask_ops import mask_losses
    compact, oc, th, wp, wn, coef = _cfg(gm, n)
    logits = torch.from_numpy(gm[f"{n}.logits"]).float().cuda().requires_grad_(True)
    masks, losses = mask_losses(logits, object_channel=oc, sharpen=mode, t_sharpen=0.25)
    (g,) = torch.autograd.grad(1.7 * losses["sharpen"], logits)
    ref = float(gm[f"{n}.sharpen.{mode}"])
    assert abs(float(losses["sharpen"]) - ref) <= 2e-5 * abs(ref) + 1e-8
    # the hinge gradient is a sign: pixels within fp32 rounding of the kink (|m_obj - max| == t) may flip
    tol = 2e-5 if mode == "kl" else 2e-2
    assert rel_l2(g.cpu().numpy(), gm[f"{n}.sharpen.{mode}.dlogits"]) < tol
    assert rel_l2(masks.detach().cpu().numpy(), gm[f"{n}.masks"]) < 1e-6


@pytest.mark.parametrize("shape", SHAPES)
def test_mask_prep_oracle_matches_reference_ops(shape):
    """entropy-only wrapper against models/rcf_model.py:433-434 and :376-378 restated with torch ops"""
    torch.manual_seed(2)
    logits = torch.randn(*shape, dtype=torch.float64) * 3
    w_mask = torch.randn(*shape, dtype=torch.float64)
    x = logits.clone().requires_grad_(True)
    all_pred_mask = F.softmax(x, dim=2)
    log_all_pred_mask = F.log_softmax(all_pred_mask, dim=2)
    ent = -(all_pred_mask * log_all_pred_mask).sum(dim=2).mean()
    ((all_pred_mask * w_mask).sum() + 0.7 * ent).backward()
    m, e = O.mask_prep_forward(logits.numpy())
    assert rel_l2(m, all_pred_mask.detach().numpy()) < 1e-13 and abs(e - float(ent)) <= 1e-13 * abs(float(ent)) + 1e-15
    assert rel_l2(O.mask_prep_backward(m, w_mask.numpy(), 0.7), x.grad.numpy()) < 1e-12


def _close(a, b, rtol, atol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return rel_l2(a, b) < rtol or np.abs(a - b).max() < atol


@pytest.mark.gpu
@pytest.mark.parametrize("n", CASES)
def test_mask_losses_kernels_match_reference_fixtures(gm, n):
    from rcf_unsupvideoseg_b200.mask_ops import mask_losses
    compact, oc, th, wp, wn, coef = _cfg(gm, n)
    logits = torch.from_numpy(gm[f"{n}.logits"]).float().cuda().requires_grad_(True)
    pl = torch.from_numpy(gm[f"{n}.pl"]).float().cuda()
    w_mask = torch.from_numpy(gm[f"{n}.w_mask"]).float().cuda()
    masks, losses = mask_losses(logits, compact_channel=compact, pl_masks=pl, object_channel=oc, pl_mask_pos_th=th,
                                pl_pos_weight=wp, pl_neg_weight=wn)
    total = (masks * w_mask).sum() + coef[0] * losses["entropy"] + coef[1] * losses["compactness"] + coef[2] * losses["pl"]
    (g,) = torch.autograd.grad(total, logits)
    assert rel_l2(masks.detach().cpu().numpy(), gm[f"{n}.masks"]) < 1e-6
    ref = gm[f"{n}.losses"]
    for i, k in enumerate(("entropy", "compactness", "pl")):
        assert abs(float(losses[k]) - ref[i]) <= 1e-5 * abs(ref[i]) + 1e-8, k
    assert rel_l2(g.cpu().numpy(), gm[f"{n}.dlogits"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + [(8, 2, 4, 96, 96), (2, 2, 4, 480, 854)])
def test_mask_prep_kernels(shape):
    from rcf_unsupvideoseg_b200.mask_ops import mask_losses, softmax_entropy
    gen = torch.Generator(device="cuda").manual_seed(4)
    logits = (torch.randn(*shape, device="cuda", generator=gen) * 3).requires_grad_(True)
    w_mask = torch.randn(*shape, device="cuda", generator=gen)
    masks, ent = softmax_entropy(logits)
    (g,) = torch.autograd.grad((masks * w_mask).sum() + 0.7 * ent, logits)
    m_o, e_o = O.mask_prep_forward(logits.detach().double().cpu().numpy())
    g_o = O.mask_prep_backward(m_o, w_mask.double().cpu().numpy(), 0.7)
    assert rel_l2(masks.detach().cpu().numpy(), m_o) < 1e-6
    assert abs(float(ent) - e_o) <= 1e-5 * abs(e_o) + 1e-7
    assert _close(g.cpu().numpy(), g_o, 1e-5, 1e-6)
    # each gradient stream alone
    (g1,) = torch.autograd.grad(softmax_entropy(logits)[1], logits)
    assert _close(g1.cpu().numpy(), O.mask_prep_backward(m_o, None, 1.0), 1e-5, 1e-9)
    (g2,) = torch.autograd.grad((softmax_entropy(logits)[0] * w_mask).sum(), logits)
    assert _close(g2.cpu().numpy(), O.mask_prep_backward(m_o, w_mask.double().cpu().numpy(), None), 1e-5, 1e-6)
    # compactness + PL at this shape against the oracle (full size: moments over 410k pixels per frame)
    B, I, K, H, W = shape
    pl = torch.rand(B, I, H, W, device="cuda", generator=gen)
    cc, oc = K - 1, 0
    masks2, losses = mask_losses(logits, compact_channel=cc, pl_masks=pl, object_channel=oc, pl_mask_pos_th=0.4,
                                 pl_pos_weight=1.5, pl_neg_weight=0.5)
    (g3,) = torch.autograd.grad(1.1 * losses["compactness"] + 0.9 * losses["pl"], logits)
    _, lo, _ = O.mask_losses_forward(logits.detach().double().cpu().numpy(), cc, pl.double().cpu().numpy(), oc, 0.4, 1.5, 0.5)
    assert torch.equal(masks2, masks)
    assert abs(float(losses["compactness"]) - lo["compactness"]) <= 2e-5 * abs(lo["compactness"]) + 1e-8
    assert abs(float(losses["pl"]) - lo["pl"]) <= 1e-5 * abs(lo["pl"]) + 1e-8
    g3_o = O.mask_losses_backward(m_o, None, None, 1.1, 0.9, cc, pl.double().cpu().numpy(), oc, 0.4, 1.5, 0.5)
    assert _close(g3.cpu().numpy(), g3_o, 2e-5, 1e-9)
    # deterministic
    assert torch.equal(ent, softmax_entropy(logits)[1])
    with pytest.raises(RuntimeError):
        softmax_entropy(torch.zeros(1, 1, 2, 4, 4))
