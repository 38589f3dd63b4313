"""Host-side behaviour of the drop-in module that needs no GPU: constructor contract, state_dict
layout, flag validation and the loud failure on CPU tensors (reference :50-148, :305-310, :324)."""
import pytest
import torch

import rcf_unsupvideoseg_b200 as pkg
from rcf_unsupvideoseg_b200 import FlowAggregationHeadWithResidual as Head


def test_state_dict_layout_matches_reference():
    h = Head(args=None, create_flownet=True, free_residual=True)
    sd = h.state_dict()
    assert list(sd.keys()) == ["flow_feat_before_agg.0.weight", "flow_feat_before_agg.0.bias",
                               "flow_feat_before_agg.2.weight", "flow_feat_before_agg.2.bias",
                               "flow_feat_after_agg.0.weight", "flow_feat_after_agg.0.bias",
                               "flow_feat_after_agg.2.weight", "flow_feat_after_agg.2.bias"]
    assert [tuple(v.shape) for v in sd.values()] == [(64, 2, 3, 3), (64,), (64, 64, 3, 3), (64,), (64, 64, 1), (64,),
                                                     (2, 64, 1), (2,)]
    assert sum(p.numel() for p in h.parameters()) == 42434
    assert h.mask_layer == 5 and h.mask_size == (48, 48) and h.pred_div_coeff == 10.


def test_golden_checkpoint_loads_strict(golden):
    h = Head(args=None, create_flownet=True, **golden.head_kwargs())
    h.load_state_dict({k: torch.from_numpy(v) for k, v in golden.params.items()}, strict=True)


def test_constructor_asserts():
    with pytest.raises(AssertionError):
        Head(args=None)                                         # create_flownet must be True (:82)
    with pytest.raises(AssertionError):
        Head(args=None, create_flownet=True, free_residual_with_affine_quadratic=True)          # :127
    with pytest.raises(AssertionError):
        Head(args=None, create_flownet=True, free_residual=True, free_residual_with_affine=True)  # :131-132


def test_cpu_tensors_fail_loudly():
    h = Head(args=None, create_flownet=True, mask_layer=2, mask_size=(4, 4), free_residual=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        h(torch.zeros(1, 2, 3, 4, 4), torch.rand(1, 2, 2, 4, 4), torch.zeros(1, 1, 2, 4, 4), torch.zeros(1, 1, 2, 4, 4),
          torch.zeros(1, 4, 4, 4), torch.zeros(1, 4, 4, 4))
    spec = pkg.LossSpec(K=2, H=4, W=4)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        pkg.rcf_motion_loss(spec, torch.rand(1, 1, 2, 4, 4), [torch.zeros(1, 2, 4, 4)], [torch.zeros(1, 4, 4, 4)],
                            thetas=[torch.zeros(1, 2, 2)])


def test_no_residual_mode_raises_like_reference():
    h = Head(args=None, create_flownet=True, mask_layer=2, mask_size=(4, 4))
    with pytest.raises(UnboundLocalError):
        h(torch.zeros(1, 2, 3, 4, 4), torch.rand(1, 2, 2, 4, 4), torch.zeros(1, 1, 2, 4, 4), torch.zeros(1, 1, 2, 4, 4),
          torch.zeros(1, 4, 4, 4), torch.zeros(1, 4, 4, 4))


def test_norm_and_clamp_flow_semantics():
    h = Head(args=None, create_flownet=True, free_residual=True, clamp_flow_t=2.0, filter_flow_t=0.5, norm_flow=False)
    f = torch.tensor([[-3.0, -0.4, 0.2, 1.0, 5.0]])
    out = h.norm_and_clamp_flow(f)
    assert out.tolist() == [[-2.0, 0.0, 0.0, 1.0, 2.0]]
    assert f.tolist() == [[-3.0, -0.4000000059604645, 0.20000000298023224, 1.0, 5.0]]   # clamp made a copy first
    h2 = Head(args=None, create_flownet=True, free_residual=True, filter_flow_t=0.5)
    g = torch.tensor([[-3.0, -0.4, 0.2]])
    h2.norm_and_clamp_flow(g)
    assert g.tolist() == [[-3.0, 0.0, 0.0]]        # reference quirk: in-place on the caller's tensor (:158-160)


def test_get_norm_flow_and_objectview():
    a = torch.ones(1, 2, 4, 8)
    h_, w_, f1, f2 = pkg.get_norm_flow(a, 2 * a)
    assert (h_, w_) == (4, 8)
    assert torch.allclose(f1[:, 0], torch.full((1, 4, 8), 0.5)) and torch.allclose(f1[:, 1], torch.full((1, 4, 8), 0.25))
    assert torch.allclose(f2, 2 * f1)
    assert list(pkg.Objectview({"a": 1}).keys()) == ["a"]


def test_loss_spec_from_flags():
    h = Head(args=None, create_flownet=True, mask_layer=4, mask_size=(8, 8), free_residual=True,
             residual_adjustment_scale=-1., clamp_flow_t=20.)
    s = h._spec(4, 8, 8, want_vis=True, vis_norm=True)
    assert s.unbounded_residual and s.D == 0 and s.Cf == 64 and s.theta_mode == 1 and s.vis_scale == (0.25, 0.25)
    h = Head(args=None, create_flownet=True, free_residual_with_affine=True, free_residual_with_affine_quadratic=True)
    assert h._spec(5, 48, 48, want_vis=False, vis_norm=False).D == 5
