"""Bilinear resize (input staging): the numpy oracle against ATen's F.interpolate -- the arithmetic the
reference calls at models/flow_aggregation_head_with_residual.py:271-273 and models/rcf_model.py:438-442 -- on CPU, and the
sm_100a kernels (csrc/rcf_resize.cu) against both on the GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import rcf_oracle as O

CASES = [  # (B, C, h, w, H, W, align_corners)
    (2, 8, 48, 48, 96, 96, False),       # DAVIS stage 1: residual 48x48 -> mask_size 96x96
    (1, 8, 24, 43, 48, 86, False),
    (2, 3, 7, 5, 13, 17, False),         # odd, non-integer scale
    (2, 3, 7, 5, 13, 17, True),
    (1, 2, 30, 40, 15, 20, False),       # downscale
    (1, 2, 31, 17, 10, 9, True),
    (1, 2, 5, 6, 1, 1, True),            # single output pixel
    (1, 2, 1, 1, 4, 6, False),           # single input pixel
    (1, 4, 16, 16, 16, 16, False),       # identity
]


def _torch_ref(x, H, W, align, g):
    xt = torch.from_numpy(x).double().requires_grad_(True)
    y = F.interpolate(xt, (H, W), mode="bilinear", align_corners=align)
    y.backward(torch.from_numpy(g).double())
    return y.detach().numpy(), xt.grad.numpy()


@pytest.mark.parametrize("case", CASES)
def test_resize_oracle_matches_aten(case):
    B, C, h, w, H, W, align = case
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B, C, h, w))
    g = rng.standard_normal((B, C, H, W))
    y_ref, gx_ref = _torch_ref(x, H, W, align, g)
    assert rel_l2(O.bilinear_resize(x, (H, W), align), y_ref) < 1e-12
    assert rel_l2(O.bilinear_resize_backward(g, (h, w), align), gx_ref) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES + [(2, 8, 240, 427, 480, 854, False)])
def test_resize_kernels_match_oracle_and_aten(case):
    from rcf_unsupvideoseg_b200.resize import resize_bilinear
    B, C, h, w, H, W, align = case
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, C, h, w, device="cuda", generator=gen).requires_grad_(True)
    g = torch.randn(B, C, H, W, device="cuda", generator=gen)
    y = resize_bilinear(x, (H, W), align)
    (gx,) = torch.autograd.grad(y, x, g)
    y_o = O.bilinear_resize(x.detach().double().cpu().numpy(), (H, W), align)
    gx_o = O.bilinear_resize_backward(g.double().cpu().numpy(), (h, w), align)
    assert rel_l2(y.detach().cpu().numpy(), y_o) < 2e-6            # fp32 vs fp64 (source positions are computed in fp32)
    assert rel_l2(gx.cpu().numpy(), gx_o) < 2e-6
    x2 = x.detach().clone().requires_grad_(True)                    # ATen on the same GPU
    y_t = F.interpolate(x2, (H, W), mode="bilinear", align_corners=align)
    (gx_t,) = torch.autograd.grad(y_t, x2, g)
    assert rel_l2(y.detach().cpu().numpy(), y_t.detach().cpu().numpy()) < 1e-6
    assert rel_l2(gx.cpu().numpy(), gx_t.cpu().numpy()) < 1e-6
    (gx_again,) = torch.autograd.grad(resize_bilinear(x, (H, W), align), x, g)
    assert torch.equal(gx, gx_again)                                # gather backward: bit-reproducible


@pytest.mark.gpu
def test_resize_two_tensors_one_launch_and_partial_grads():
    from rcf_unsupvideoseg_b200.resize import resize_bilinear, resize_bilinear_multi
    gen = torch.Generator(device="cuda").manual_seed(6)
    a = torch.randn(2, 8, 12, 10, device="cuda", generator=gen).requires_grad_(True)
    b = torch.randn(2, 8, 12, 10, device="cuda", generator=gen)        # no grad
    ya, yb = resize_bilinear_multi([a, b], (24, 20))
    assert torch.equal(ya, resize_bilinear(a, (24, 20))) and torch.equal(yb, resize_bilinear(b, (24, 20)))
    (ga,) = torch.autograd.grad(ya.sum() + yb.sum(), a)
    assert torch.allclose(ga, 4 * torch.ones_like(ga), atol=1e-5)        # 2x upsampling: every input pixel carries total weight 2*2
    with pytest.raises(RuntimeError):
        resize_bilinear(torch.zeros(1, 1, 4, 4), (8, 8))                 # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(3, 2, 30, 53, 12, 12, False, None), (2, 2, 48, 85, 96, 96, True, (1.5, 0.75)),
                                  (1, 2, 480, 854, 96, 96, False, (0.2, 0.2))])
def test_stage_flow_hwc_matches_reference_op_sequence(case):
    """scale (FlowTransform.scale_flow) + np.transpose(flow, (2,0,1)) (dataset/transforms.py:842-850) + bilinear resize
    (models/rcf_model.py:438-442) against the oracle and against the same sequence in ATen."""
    from rcf_unsupvideoseg_b200.resize import stage_flow_hwc
    N, C_, h, w, H, W, align, cs = case
    gen = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(N, h, w, C_, device="cuda", generator=gen) * 6
    y = stage_flow_hwc(x, (H, W), align, cs)
    ref_in = x.double().cpu().numpy()
    if cs is not None:
        ref_in = ref_in * np.asarray(cs)
    y_o = O.bilinear_resize(np.transpose(ref_in, (0, 3, 1, 2)), (H, W), align)
    # the source positions are computed in fp32 as ATen does; at column ~850 an fp32 ulp is 6e-5 pixels and the input is
    # white noise, so the fp64 oracle is matched to ~1e-4 there (1e-6 where the positions are small)
    tol = 1e-3 if max(h, w) > 256 else 1e-5
    assert y.shape == (N, C_, H, W) and rel_l2(y.cpu().numpy(), y_o) < tol
    xt = x * torch.tensor(cs, device="cuda") if cs is not None else x
    y_t = F.interpolate(xt.permute(0, 3, 1, 2).contiguous(), (H, W), mode="bilinear", align_corners=align)
    assert rel_l2(y.cpu().numpy(), y_t.cpu().numpy()) < (1e-4 if max(h, w) > 256 else 1e-6)   # FMA contraction of the position
