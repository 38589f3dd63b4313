"""Pin the numpy oracle against the reference's own outputs (committed fixtures).

Every fixture was produced by the unmodified reference module (tests/golden/make_golden.py).
The oracle (fp64) must reproduce the reference's fp64 run to ~1e-9 everywhere: loss scalars,
gradients of masks / residual maps / all 8 parameters, and every returned visualisation flow.
"""
import numpy as np
import pytest

from conftest import rel_l2
from oracle import rcf_oracle as O


def oracle_cfg(g):
    kw = dict(g.head_kwargs())
    return O.OracleConfig(**kw)


def test_oracle_matches_reference_fp64(golden):
    cfg = oracle_cfg(golden)
    params = {k: v.astype(np.float64) for k, v in golden.params.items()}
    masks, fw, bw, rfw, rbw = [a.astype(np.float64) for a in golden.inputs]
    flows, loss, caches = O.head_forward(masks, fw, bw, rfw, rbw, params, cfg)
    for k in ("seg_fw", "seg_bw", "seg"):
        assert loss[k] == pytest.approx(float(golden.ref("f64", "loss." + k)), rel=1e-11)
    for k, v in flows.items():
        if v:
            assert rel_l2(v[0], golden.ref("f64", "flows." + k)) < 1e-10, k
        else:
            assert not golden.has("f64", "flows." + k)
    grads = O.head_backward(caches, params, gbar=golden.gbar)
    tol = 1e-6 if cfg.D == 5 else 1e-9     # raw-coordinate quadratic normal equations: cond ~1e9 even in fp64
    for k in ("d_masks", "d_resid_fw", "d_resid_bw"):
        assert rel_l2(grads[k], golden.ref("f64", k)) < tol, k
    for k, v in grads["params"].items():
        assert rel_l2(v, golden.ref("f64", "dparam." + k)) < max(tol, 1e-8), k


def test_centred_basis_is_equivalent(golden):
    """The CUDA kernels fit in a centred, scaled coordinate basis; the fit is basis-invariant."""
    cfg = oracle_cfg(golden)
    if cfg.D == 0:
        pytest.skip("no affine fit in this mode")
    params = {k: v.astype(np.float64) for k, v in golden.params.items()}
    ins = [a.astype(np.float64) for a in golden.inputs]
    _, loss_r, c_r = O.head_forward(*ins, params, cfg, basis="reference")
    _, loss_c, c_c = O.head_forward(*ins, params, cfg, basis="centred")
    assert loss_c["seg"] == pytest.approx(loss_r["seg"], rel=1e-9)
    g_r = O.head_backward(c_r, params, 1.0)
    g_c = O.head_backward(c_c, params, 1.0)
    tol = 1e-5 if cfg.D == 5 else 1e-9
    for k in ("d_masks", "d_resid_fw", "d_resid_bw"):
        assert rel_l2(g_c[k], g_r[k]) < tol, k


def test_reference_fp32_noise_floor_recorded(golden):
    """Documents how far the reference's own fp32 run is from its fp64 run (the parity floor)."""
    loss32 = float(golden.ref("f32", "loss.seg"))
    loss64 = float(golden.ref("f64", "loss.seg"))
    assert abs(loss32 - loss64) / abs(loss64) < 1e-5
    floor = rel_l2(golden.ref("f32", "d_masks"), golden.ref("f64", "d_masks"))
    D5 = golden.kwargs.get("free_residual_with_affine_quadratic", False)
    assert floor < (0.2 if D5 else 5e-3)


def test_no_residual_mode_is_rejected():
    cfg = O.OracleConfig(mask_layer=2, mask_size=(4, 4), num_flow_feat_channels=2)
    with pytest.raises(AssertionError):
        O.direction_forward(np.ones((1, 2, 4, 4)), np.zeros((1, 2, 4, 4)), np.zeros((1, 4, 4, 4)),
                            O.init_params(cfg), cfg)


def test_bilinear_resize_backward_is_adjoint():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 3, 5, 7))
    y = rng.normal(size=(2, 3, 11, 13))
    lhs = np.sum(O.bilinear_resize(x, (11, 13)) * y)
    rhs = np.sum(x * O.bilinear_resize_backward(y, (5, 7)))
    assert lhs == pytest.approx(rhs, rel=1e-12)


def test_conv_backward_is_adjoint():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(2, 3, 6, 5))
    w = rng.normal(size=(4, 3, 3, 3))
    b = rng.normal(size=(4,))
    dout = rng.normal(size=(2, 4, 6, 5))
    dx, dw, db = O.conv2d_same_backward(x, w, dout)
    eps = 1e-6
    dxn = rng.normal(size=x.shape)
    dwn = rng.normal(size=w.shape)
    num_x = (np.sum(O.conv2d_same(x + eps * dxn, w, b) * dout) - np.sum(O.conv2d_same(x - eps * dxn, w, b) * dout)) / (2 * eps)
    num_w = (np.sum(O.conv2d_same(x, w + eps * dwn, b) * dout) - np.sum(O.conv2d_same(x, w - eps * dwn, b) * dout)) / (2 * eps)
    assert np.sum(dx * dxn) == pytest.approx(num_x, rel=1e-7)
    assert np.sum(dw * dwn) == pytest.approx(num_w, rel=1e-7)
    assert np.allclose(db, dout.sum(axis=(0, 2, 3)))
