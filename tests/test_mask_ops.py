"""Caller-side mask preparation: oracle vs the reference's own op sequence (CPU), kernels vs both (GPU)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import rcf_oracle as O

SHAPES = [(2, 2, 4, 12, 16), (1, 2, 3, 7, 9), (2, 2, 8, 6, 10), (1, 1, 1, 4, 4), (3, 2, 5, 5, 5)]


def _reference_ops(logits, w_mask, w_ent):
    """models/rcf_model.py:433-434 and :376-378 verbatim in semantics; returns masks, entropy, d(w_mask.masks + w_ent*entropy)/dlogits."""
    x = logits.clone().requires_grad_(True)
    all_pred_mask = F.softmax(x, dim=2)
    log_all_pred_mask = F.log_softmax(all_pred_mask, dim=2)
    ent = -(all_pred_mask * log_all_pred_mask).sum(dim=2).mean()
    ((all_pred_mask * w_mask).sum() + w_ent * ent).backward()
    return all_pred_mask.detach(), ent.detach(), x.grad


@pytest.mark.parametrize("shape", SHAPES)
def test_mask_prep_oracle_matches_reference_ops(shape):
    torch.manual_seed(2)
    logits = torch.randn(*shape, dtype=torch.float64) * 3
    w_mask = torch.randn(*shape, dtype=torch.float64)
    m_ref, e_ref, g_ref = _reference_ops(logits, w_mask, 0.7)
    m, e = O.mask_prep_forward(logits.numpy())
    assert rel_l2(m, m_ref.numpy()) < 1e-13 and abs(e - float(e_ref)) <= 1e-13 * abs(float(e_ref)) + 1e-15
    g = O.mask_prep_backward(m, w_mask.numpy(), 0.7)
    assert rel_l2(g, g_ref.numpy()) < 1e-12
    g0 = O.mask_prep_backward(m, None, 1.0)                       # entropy only
    _, _, g0_ref = _reference_ops(logits, torch.zeros_like(w_mask), 1.0)
    assert rel_l2(g0, g0_ref.numpy()) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES + [(8, 2, 4, 96, 96), (2, 2, 4, 480, 854)])
def test_mask_prep_kernels(shape):
    from rcf_unsupvideoseg_b200.mask_ops import softmax_entropy
    gen = torch.Generator(device="cuda").manual_seed(4)
    logits = (torch.randn(*shape, device="cuda", generator=gen) * 3).requires_grad_(True)
    w_mask = torch.randn(*shape, device="cuda", generator=gen)
    masks, ent = softmax_entropy(logits)
    (g,) = torch.autograd.grad((masks * w_mask).sum() + 0.7 * ent, logits)
    m_o, e_o = O.mask_prep_forward(logits.detach().double().cpu().numpy())
    g_o = O.mask_prep_backward(m_o, w_mask.double().cpu().numpy(), 0.7)
    assert rel_l2(masks.detach().cpu().numpy(), m_o) < 1e-6
    assert abs(float(ent) - e_o) <= 1e-5 * abs(e_o) + 1e-7
    assert rel_l2(g.cpu().numpy(), g_o) < 1e-5 or np.abs(g.cpu().numpy() - g_o).max() < 1e-6
    # each gradient stream alone
    (g1,) = torch.autograd.grad(softmax_entropy(logits)[1], logits)
    g1_o = O.mask_prep_backward(m_o, None, 1.0)
    assert rel_l2(g1.cpu().numpy(), g1_o) < 1e-5 or np.abs(g1.cpu().numpy() - g1_o).max() < 1e-9
    (g2,) = torch.autograd.grad((softmax_entropy(logits)[0] * w_mask).sum(), logits)
    g2_o = O.mask_prep_backward(m_o, w_mask.double().cpu().numpy(), None)
    assert rel_l2(g2.cpu().numpy(), g2_o) < 1e-5 or np.abs(g2.cpu().numpy() - g2_o).max() < 1e-6
    # deterministic
    assert torch.equal(ent, softmax_entropy(logits)[1])
    with pytest.raises(RuntimeError):
        softmax_entropy(torch.zeros(1, 1, 2, 4, 4))
