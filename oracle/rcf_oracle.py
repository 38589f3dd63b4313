"""CPU oracle for the RCF relaxed-common-fate motion loss (numpy, fp64).

TEST INFRASTRUCTURE ONLY.  Nothing under ``rcf_unsupvideoseg_b200/`` may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline leg use it, and only
as the checker.  The shipped path is the sm_100a CUDA library behind ``include/rcf_loss.h``.

What it restates (all citations into /root/reference/):
  * flow preparation ............ models/flow_aggregation_head_with_residual.py:150-162
  * mask normalise + pooling .... :242-256
  * segment MLP ................. :95-101, :258
  * piece-wise constant flow .... :260-265
  * weighted LSQ affine fit ..... :164-233  (D=2, or D=5 when "quadratic")
  * residual (tanh / unbounded) . :268-304  (bilinear resize :271-273 / :294-296)
  * loss (L1 / robust) .......... :359-368
  * visualisation flows ......... :18-30, :370-395
  * backward .................... closed form, SURVEY.md 8(a)-math (the reference relies on autograd)

Pinning: ``tests/golden/*.npz`` were produced by importing the unmodified reference module in the
build container (``tests/golden/make_golden.py``); ``tests/test_oracle_golden.py`` checks every
function here against them (loss, all input gradients, all 8 parameter gradients, all returned
flows).  The reference repo ships no tests or golden vectors of its own for this path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

LEAK = 0.1  # nn.LeakyReLU(0.1), reference :88,:92,:98


@dataclass
class OracleConfig:
    """Mirror of the reference constructor's keyword arguments (reference :50-74)."""
    mask_layer: int = 5
    flow_feat_before_agg_kernel_size: int = 3
    num_flow_feat_channels: int = 64
    outlier_robust_loss: bool = False
    eps: float = 0.01
    q: float = 0.4
    mask_size: Tuple[int, int] = (48, 48)
    residual_adjustment_scale: float = 10.0
    norm_flow: bool = False
    clamp_flow_t: Optional[float] = None
    filter_flow_t: Optional[float] = None
    free_residual: bool = False
    free_residual_with_affine: bool = False
    free_residual_with_affine_quadratic: bool = False
    allow_residual_resize: bool = False
    pred_div_coeff: float = 10.0

    @property
    def D(self) -> int:
        if not self.free_residual_with_affine:
            return 0
        return 5 if self.free_residual_with_affine_quadratic else 2


# ----------------------------------------------------------------------------------------------
# small dense-layer helpers (the learned part of the head; reference :84-101)
# ----------------------------------------------------------------------------------------------

def lrelu(x):
    return np.where(x >= 0, x, LEAK * x)


def lrelu_grad(y_or_pre):
    # LeakyReLU with a positive slope preserves sign, so the mask can be taken on either side.
    return np.where(y_or_pre >= 0, 1.0, LEAK)


def conv2d_same(x, w, b):
    """Cross-correlation with zero padding (k-1)//2, stride 1 (nn.Conv2d semantics)."""
    k = w.shape[-1]
    pad = (k - 1) // 2
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    win = sliding_window_view(xp, (k, k), axis=(2, 3))       # [B,Cin,H',W',k,k]
    out = np.einsum("bihwyx,oiyx->bohw", win, w, optimize=True)
    return out + b[None, :, None, None]


def conv2d_same_backward(x, w, dout, need_dx=True):
    k = w.shape[-1]
    pad = (k - 1) // 2
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    win = sliding_window_view(xp, (k, k), axis=(2, 3))
    # output spatial size may be smaller than input for even k; crop windows accordingly
    dw = np.einsum("bihwyx,bohw->oiyx", win, dout, optimize=True)
    db = dout.sum(axis=(0, 2, 3))
    dx = None
    if need_dx:
        H, W = x.shape[2:]
        Ho, Wo = dout.shape[2:]
        dxp = np.zeros_like(xp)
        for yy in range(k):
            for xx in range(k):
                dxp[:, :, yy:yy + Ho, xx:xx + Wo] += np.einsum("bohw,oi->bihw", dout, w[:, :, yy, xx], optimize=True)
        dx = dxp[:, :, pad:pad + H, pad:pad + W]
    return dx, dw, db


def _interp_axis(n_in: int, n_out: int, align_corners: bool = False):
    """Source indices / weights of F.interpolate(mode='bilinear') (ATen upsample_bilinear2d: scale = in/out and
    src = max(scale*(dst+0.5)-0.5, 0), or scale = (in-1)/(out-1) and src = scale*dst with align_corners)."""
    if align_corners:
        scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        src = np.arange(n_out, dtype=np.float64) * scale
    else:
        scale = n_in / n_out
        src = (np.arange(n_out, dtype=np.float64) + 0.5) * scale - 0.5
    src = np.maximum(src, 0.0)
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = src - i0
    return i0, i1, lam


def bilinear_resize(x, size, align_corners: bool = False):
    """[B,C,h,w] -> [B,C,H,W]; reference :271-273 (F.interpolate defaults) and, with align_corners, the caller's
    mmseg resize of the flows (models/rcf_model.py:438-442)."""
    H, W = size
    h, w = x.shape[2:]
    y0, y1, ly = _interp_axis(h, H, align_corners)
    x0, x1, lx = _interp_axis(w, W, align_corners)
    rows = x[:, :, y0, :] * (1 - ly)[None, None, :, None] + x[:, :, y1, :] * ly[None, None, :, None]
    return rows[:, :, :, x0] * (1 - lx) + rows[:, :, :, x1] * lx


def bilinear_resize_backward(dout, in_size, align_corners: bool = False):
    h, w = in_size
    B, C, H, W = dout.shape
    y0, y1, ly = _interp_axis(h, H, align_corners)
    x0, x1, lx = _interp_axis(w, W, align_corners)
    drows = np.zeros((B, C, H, w), dtype=dout.dtype)
    np.add.at(drows, (slice(None), slice(None), slice(None), x0), dout * (1 - lx))
    np.add.at(drows, (slice(None), slice(None), slice(None), x1), dout * lx)
    dx = np.zeros((B, C, h, w), dtype=dout.dtype)
    np.add.at(dx, (slice(None), slice(None), y0, slice(None)), drows * (1 - ly)[None, None, :, None])
    np.add.at(dx, (slice(None), slice(None), y1, slice(None)), drows * ly[None, None, :, None])
    return dx


# ----------------------------------------------------------------------------------------------
# path pieces
# ----------------------------------------------------------------------------------------------

def prepare_flow(flow, cfg: OracleConfig):
    """reference :150-162 (returns a new array; the reference's in-place filter quirk is a caller-
    visible side effect, not part of the value)."""
    flow = np.array(flow, dtype=np.float64, copy=True)
    if cfg.norm_flow:
        flow = flow / np.abs(flow).max()
    if cfg.clamp_flow_t is not None:
        flow = np.clip(flow, -cfg.clamp_flow_t, cfg.clamp_flow_t)
    if cfg.filter_flow_t is not None:
        flow[np.abs(flow) < cfg.filter_flow_t] = 0.0
    return flow


def coord_features(H: int, W: int, D: int, basis: str = "reference"):
    """[P, D] coordinate features.  'reference' = raw (row, col[, row^2, col^2, row*col]) exactly
    as the reference builds them (:135-148).  'centred' = the same function space in a
    well-conditioned basis (coordinates shifted to the frame centre and scaled to [-1, 1]); the
    de-meaned weighted least-squares fit is invariant under this change of basis, which is what the
    CUDA kernels use."""
    r, c = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    if basis == "centred":
        cy, cx = (H - 1) / 2.0, (W - 1) / 2.0
        r = (r - cy) / max(cy, 1.0)
        c = (c - cx) / max(cx, 1.0)
    elif basis != "reference":
        raise ValueError(basis)
    feats = [r, c]
    if D == 5:
        feats += [r * r, c * c, r * c]
    elif D != 2:
        raise ValueError(D)
    return np.stack(feats, axis=-1).reshape(H * W, D)


@dataclass
class DirCache:
    """Everything one direction's backward needs."""
    cfg: OracleConfig
    M: np.ndarray          # [B,K,P]
    F: np.ndarray          # [B,2,P] prepared flow
    F_img: np.ndarray      # [B,2,H,W]
    R_in_shape: Tuple[int, ...]
    T: np.ndarray          # [B,2,K,P]  tanh(R/tau) or R when scale == -1
    S: np.ndarray          # [B,K]
    Mn: np.ndarray         # [B,K,P]
    act1: np.ndarray       # conv1 output after lrelu [B,Cf,H,W]
    G: np.ndarray          # [B,Cf,P]
    pool: np.ndarray       # [B,Cf,K]
    hpre: np.ndarray       # [B,Cf,K]
    theta: np.ndarray      # [B,2,K]
    U: Optional[np.ndarray] = None      # [P,D]
    mu_u: Optional[np.ndarray] = None   # [B,K,D]
    mu_F: Optional[np.ndarray] = None   # [B,K,2]
    SFu: Optional[np.ndarray] = None    # [B,K,2,D]
    Suu: Optional[np.ndarray] = None    # [B,K,D,D]
    A: Optional[np.ndarray] = None      # [B,K,2,D]
    pred: np.ndarray = None
    agg: np.ndarray = None
    res: np.ndarray = None
    aff: Optional[np.ndarray] = None
    loss: float = 0.0
    resized: bool = False


def direction_forward(mask, flow, resid, params: Dict[str, np.ndarray], cfg: OracleConfig,
                      basis: str = "reference") -> DirCache:
    """One call of aggregate_flow_with_residual (:235-310) plus its loss term (:359-368).

    mask [B,K,H,W]; flow [B,2,H,W] (raw, un-prepared); resid [B,2K,h,w];
    params: the 8 state_dict tensors keyed as in the reference.
    """
    assert cfg.free_residual or cfg.free_residual_with_affine, \
        "the reference raises UnboundLocalError without a residual mode (:305-310)"
    mask = np.asarray(mask, dtype=np.float64)
    resid = np.asarray(resid, dtype=np.float64)
    B, K, H, W = mask.shape
    P = H * W
    F_img = prepare_flow(flow, cfg)

    M = mask.reshape(B, K, P)
    S = M.sum(axis=2)                                   # :242-243
    Mn = M / S[:, :, None]

    w0, b0 = params["flow_feat_before_agg.0.weight"], params["flow_feat_before_agg.0.bias"]
    w2, b2 = params["flow_feat_before_agg.2.weight"], params["flow_feat_before_agg.2.bias"]
    act1 = lrelu(conv2d_same(F_img, w0, b0))            # :246
    G_img = lrelu(conv2d_same(act1, w2, b2))
    assert G_img.shape[2:] == (H, W), "conv features / mask spatial mismatch (:247-248)"
    Cf = G_img.shape[1]
    G = G_img.reshape(B, Cf, P)
    pool = np.einsum("bfp,bkp->bfk", G, Mn, optimize=True)   # :251-256

    m1w, m1b = params["flow_feat_after_agg.0.weight"][:, :, 0], params["flow_feat_after_agg.0.bias"]
    m2w, m2b = params["flow_feat_after_agg.2.weight"][:, :, 0], params["flow_feat_after_agg.2.bias"]
    hpre = np.einsum("gf,bfk->bgk", m1w, pool) + m1b[None, :, None]     # :258
    theta = np.einsum("cg,bgk->bck", m2w, lrelu(hpre)) + m2b[None, :, None]

    Fp = F_img.reshape(B, 2, P)
    agg = np.einsum("bck,bkp->bcp", theta, M)           # :260-265

    cache = DirCache(cfg=cfg, M=M, F=Fp, F_img=F_img, R_in_shape=resid.shape, T=None, S=S, Mn=Mn,
                     act1=act1, G=G, pool=pool, hpre=hpre, theta=theta)

    aff = None
    if cfg.free_residual_with_affine:                   # :164-233
        D = cfg.D
        U = coord_features(H, W, D, basis)
        mu_F = np.einsum("bkp,bcp->bkc", Mn, Fp)
        mu_u = np.einsum("bkp,pd->bkd", Mn, U)
        f = Fp[:, None, :, :] - mu_F[:, :, :, None]                      # [B,K,2,P]
        v = U.T[None, None, :, :] - mu_u[:, :, :, None]                  # [B,K,D,P]
        SFu = np.einsum("bkp,bkcp,bkdp->bkcd", Mn, f, v, optimize=True)
        Suu = np.einsum("bkp,bkdp,bkep->bkde", Mn, v, v, optimize=True)
        A = np.linalg.solve(Suu, np.swapaxes(SFu, 2, 3))                 # Suu X = SFu^T
        A = np.swapaxes(A, 2, 3)                                         # [B,K,2,D]
        affk = np.einsum("bkcd,bkdp->bkcp", A, v)
        aff = np.einsum("bkp,bkcp->bcp", M, affk)
        cache.U, cache.mu_u, cache.mu_F, cache.SFu, cache.Suu, cache.A = U, mu_u, mu_F, SFu, Suu, A

    # residual :268-304
    if cfg.allow_residual_resize and tuple(resid.shape[-2:]) != tuple(cfg.mask_size):
        resid_full = bilinear_resize(resid, cfg.mask_size)
        cache.resized = True
    else:
        resid_full = resid
    assert resid_full.shape[-2:] == (H, W)
    R = resid_full.reshape(B, 2, K, P)                  # channel index c*K + k (:274-275)
    s = cfg.residual_adjustment_scale
    if cfg.free_residual and s == -1.0:                 # :282-286
        T = R
        res = np.einsum("bckp,bkp->bcp", T, M)
    else:
        T = np.tanh(R / cfg.pred_div_coeff)
        res = np.einsum("bckp,bkp->bcp", T, M) * s
    cache.T = T

    pred = agg + res + (aff if aff is not None else 0.0)
    d = Fp - pred
    if cfg.outlier_robust_loss:
        loss = np.mean((np.abs(d) + cfg.eps) ** cfg.q)
    else:
        loss = np.mean(np.abs(d))
    cache.pred, cache.agg, cache.res, cache.aff, cache.loss = pred, agg, res, aff, float(loss)
    return cache


def direction_backward(c: DirCache, params: Dict[str, np.ndarray], gbar: float = 1.0):
    """Closed-form backward of one direction.

    Returns dict with d_mask [B,K,H,W], d_resid (input resolution), parameter grads (same keys as
    params) and d_theta [B,2,K] / d_pool [B,Cf,K] (useful when testing the kernels in isolation).
    """
    cfg = c.cfg
    B, K, P = c.M.shape
    H, W = c.F_img.shape[2:]
    N = c.F.size                                        # B*2*P elements in the mean
    d = c.F - c.pred
    if cfg.outlier_robust_loss:
        psi = cfg.q * (np.abs(d) + cfg.eps) ** (cfg.q - 1.0) * np.sign(d)
    else:
        psi = np.sign(d)
    g = -gbar * psi / N                                 # dl/dpred  [B,2,P]

    s = cfg.residual_adjustment_scale
    tanh_mode = not (cfg.free_residual and s == -1.0)
    if tanh_mode:
        dR = g[:, :, None, :] * (s / cfg.pred_div_coeff) * (1.0 - c.T ** 2) * c.M[:, None, :, :]
        sT = s * c.T
    else:
        dR = g[:, :, None, :] * c.M[:, None, :, :]
        sT = c.T

    theta_bar = np.einsum("bcp,bkp->bck", g, c.M)       # [B,2,K]

    # direct term: sum_c g_cp * (theta_ck + A_kc.v + s T_ckp)
    q = c.theta[:, :, :, None] + sT                     # [B,2,K,P]
    nbar = np.zeros((B, K, P))
    corr = np.zeros((B, K))
    if cfg.free_residual_with_affine:
        v = c.U.T[None, None, :, :] - c.mu_u[:, :, :, None]          # [B,K,D,P]
        f = c.F[:, None, :, :] - c.mu_F[:, :, :, None]               # [B,K,2,P]
        q = q + np.einsum("bkcd,bkdp->bckp", c.A, v)
        A_bar = np.einsum("bcp,bkp,bkdp->bkcd", g, c.M, v, optimize=True)
        Suu_inv = np.linalg.inv(c.Suu)
        SFu_bar = A_bar @ Suu_inv                                    # [B,K,2,D]
        Suu_bar = -np.swapaxes(c.A, 2, 3) @ A_bar @ Suu_inv          # [B,K,D,D]
        mu_u_bar = -np.einsum("bkcd,bck->bkd", c.A, theta_bar)
        nbar += np.einsum("bkcd,bkcp,bkdp->bkp", SFu_bar, f, v, optimize=True)
        nbar += np.einsum("bkde,bkdp,bkep->bkp", Suu_bar, v, v, optimize=True)
        nbar += np.einsum("bkd,pd->bkp", mu_u_bar, c.U)
        corr += np.einsum("bkcd,bkcd->bk", SFu_bar, c.SFu) + np.einsum("bkde,bkde->bk", Suu_bar, c.Suu) \
            + np.einsum("bkd,bkd->bk", mu_u_bar, c.mu_u)
    dM = np.einsum("bcp,bckp->bkp", g, q)

    # segment MLP backward (:95-101)
    m1w = params["flow_feat_after_agg.0.weight"][:, :, 0]
    m2w = params["flow_feat_after_agg.2.weight"][:, :, 0]
    h = lrelu(c.hpre)
    grads: Dict[str, np.ndarray] = {}
    grads["flow_feat_after_agg.2.weight"] = np.einsum("bck,bgk->cg", theta_bar, h)[:, :, None]
    grads["flow_feat_after_agg.2.bias"] = theta_bar.sum(axis=(0, 2))
    dh = np.einsum("cg,bck->bgk", m2w, theta_bar) * lrelu_grad(c.hpre)
    grads["flow_feat_after_agg.0.weight"] = np.einsum("bgk,bfk->gf", dh, c.pool)[:, :, None]
    grads["flow_feat_after_agg.0.bias"] = dh.sum(axis=(0, 2))
    pool_bar = np.einsum("gf,bgk->bfk", m1w, dh)         # [B,Cf,K]

    nbar += np.einsum("bfk,bfp->bkp", pool_bar, c.G)
    corr += np.einsum("bfk,bfk->bk", pool_bar, c.pool)
    dM = dM + (nbar - corr[:, :, None]) / c.S[:, :, None]

    # conv feature branch backward (:84-93); no gradient to the (ground-truth) flow is needed
    dG = np.einsum("bfk,bkp->bfp", pool_bar, c.Mn).reshape(B, -1, H, W)
    G_img = c.G.reshape(B, -1, H, W)
    dpre2 = dG * lrelu_grad(G_img)
    dact1, dw2, db2 = conv2d_same_backward(c.act1, params["flow_feat_before_agg.2.weight"], dpre2)
    dpre1 = dact1 * lrelu_grad(c.act1)
    _, dw0, db0 = conv2d_same_backward(c.F_img, params["flow_feat_before_agg.0.weight"], dpre1, need_dx=False)
    grads["flow_feat_before_agg.2.weight"], grads["flow_feat_before_agg.2.bias"] = dw2, db2
    grads["flow_feat_before_agg.0.weight"], grads["flow_feat_before_agg.0.bias"] = dw0, db0

    dR = dR.reshape(B, 2 * K, H, W)
    if c.resized:
        dR = bilinear_resize_backward(dR, c.R_in_shape[-2:])
    return {"d_mask": dM.reshape(B, K, H, W), "d_resid": dR, "params": grads,
            "d_theta": theta_bar, "d_pool": pool_bar, "d_feat": dG}


def norm_flow_for_vis(x):
    """get_norm_flow (:18-30): channel 0 / (H/2), channel 1 / (W/2)."""
    H, W = x.shape[2:]
    return np.concatenate([x[:, 0:1] / (H / 2.0), x[:, 1:2] / (W / 2.0)], axis=1)


def head_forward(masks, gt_fw_flows, gt_bw_flows, resid_fw, resid_bw, params, cfg: OracleConfig,
                 basis: str = "reference"):
    """FlowAggregationHeadWithResidual.forward (:312-399).  masks [B,2,K,H,W]; flows [B,1,2,H,W]."""
    masks = np.asarray(masks, dtype=np.float64)
    assert masks.shape[1] == 2, "Other im_num not implemented (:324)"
    B, _, K, H, W = masks.shape
    fw = direction_forward(masks[:, 0], np.asarray(gt_fw_flows)[:, 0], resid_fw, params, cfg, basis)
    bw = direction_forward(masks[:, 1], np.asarray(gt_bw_flows)[:, 0], resid_bw, params, cfg, basis)

    def vis(a, b):
        return np.concatenate([norm_flow_for_vis(a.reshape(B, 2, H, W)),
                               norm_flow_for_vis(b.reshape(B, 2, H, W))], axis=1)

    flows = {"gt_flow": [vis(fw.F, bw.F)], "pred_flow": [vis(fw.pred, bw.pred)],
             "agg_flow": [vis(fw.agg, bw.agg)], "residual_adj": [vis(fw.res, bw.res)],
             "affine_flow": [vis(fw.aff, bw.aff)] if fw.aff is not None else []}
    loss = {"seg_fw": fw.loss, "seg_bw": bw.loss, "seg": fw.loss + bw.loss}
    return flows, loss, (fw, bw)


def head_backward(caches, params, gbar: float = 1.0):
    """Gradient of loss['seg'] * gbar w.r.t. masks, both residual maps and the 8 parameters."""
    fw, bw = caches
    gf = direction_backward(fw, params, gbar)
    gb = direction_backward(bw, params, gbar)
    d_masks = np.stack([gf["d_mask"], gb["d_mask"]], axis=1)
    pg = {k: gf["params"][k] + gb["params"][k] for k in gf["params"]}
    return {"d_masks": d_masks, "d_resid_fw": gf["d_resid"], "d_resid_bw": gb["d_resid"], "params": pg,
            "fw": gf, "bw": gb}


def init_params(cfg: OracleConfig, seed: int = 1, scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Deterministic stand-in parameters (uniform +-1/sqrt(fan_in), like nn.Conv default bounds)."""
    rng = np.random.default_rng(seed)
    Cf, k = cfg.num_flow_feat_channels, cfg.flow_feat_before_agg_kernel_size

    def u(shape, fan_in):
        b = scale / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape)

    return {
        "flow_feat_before_agg.0.weight": u((Cf, 2, k, k), 2 * k * k),
        "flow_feat_before_agg.0.bias": u((Cf,), 2 * k * k),
        "flow_feat_before_agg.2.weight": u((Cf, Cf, k, k), Cf * k * k),
        "flow_feat_before_agg.2.bias": u((Cf,), Cf * k * k),
        "flow_feat_after_agg.0.weight": u((Cf, Cf, 1), Cf),
        "flow_feat_after_agg.0.bias": u((Cf,), Cf),
        "flow_feat_after_agg.2.weight": u((2, Cf, 1), Cf),
        "flow_feat_after_agg.2.bias": u((2,), Cf),
    }


def synthetic_inputs(B, K, H, W, seed=0, resid_hw=None, mask_sharpness=1.0, dtype=np.float32):
    """SURVEY.md 8(d) synthetic inputs: softmax(N(0,2^2)) masks, N(0,8^2) flows, N(0,5^2) residuals."""
    rng = np.random.default_rng(seed)
    logits = rng.normal(0.0, 2.0, size=(B, 2, K, H, W)) * mask_sharpness
    logits -= logits.max(axis=2, keepdims=True)
    e = np.exp(logits)
    masks = e / e.sum(axis=2, keepdims=True)
    fw = rng.normal(0.0, 8.0, size=(B, 1, 2, H, W))
    bw = rng.normal(0.0, 8.0, size=(B, 1, 2, H, W))
    rh, rw = resid_hw if resid_hw is not None else (H, W)
    rfw = rng.normal(0.0, 5.0, size=(B, 2 * K, rh, rw))
    rbw = rng.normal(0.0, 5.0, size=(B, 2 * K, rh, rw))
    return tuple(a.astype(dtype) for a in (masks, fw, bw, rfw, rbw))


# ----------------------------------------------------------------------------------------------
# caller-side mask preparation and mask losses -- SURVEY 8f rank 2
#   models/rcf_model.py:433-434 (softmax, log-softmax of the result), :376-378 (entropy), :380-408 (PL / CRF loss),
#   models/compactness_head.py:33-56 (compactness)
# ----------------------------------------------------------------------------------------------
def _pl_target(target, th):
    return target if th == -1 else (target > th).astype(target.dtype)


def _sharpen(p, T, axis):
    """utils/loss_utils.py:105-108"""
    sp = p ** (1.0 / T)
    return sp / sp.sum(axis=axis, keepdims=True)


def mask_losses_forward(logits, compact_channel=None, target=None, object_channel=None, pl_th=-1.0, wpos=1.0, wneg=1.0,
                        sharpen=None, t_sharpen=0.25):
    """logits [B,I,K,H,W] -> (masks, dict(entropy, compactness, pl, sharpen), centres [B*I,2]).
    sharpen: None | 'kl' | 'object_hinge' (models/rcf_model.py:350-374)."""
    x = logits - logits.max(axis=2, keepdims=True)
    e = np.exp(x)
    m = e / e.sum(axis=2, keepdims=True)
    lse = np.log(np.exp(m).sum(axis=2, keepdims=True))
    ls = m - lse                                                     # :434 log_softmax OF the masks
    out = {"entropy": float(-(m * ls).sum(axis=2).mean()), "compactness": 0.0, "pl": 0.0, "sharpen": 0.0}       # :376-378
    if sharpen == 'kl':                                              # :370-373; F.kl_div(input, target) = target*(log target - input)
        t = _sharpen(m, t_sharpen, 2)
        with np.errstate(divide='ignore', invalid='ignore'):
            kl = np.where(t > 0, t * (np.log(t) - ls), 0.0)
        out["sharpen"] = float(kl.mean())
    elif sharpen == 'object_hinge':                                  # :362-369
        other = m.copy()
        other[:, :, object_channel] = 0.0
        diff = np.abs(m[:, :, object_channel] - other.max(axis=2))
        out["sharpen"] = float(np.maximum(t_sharpen - diff, 0.0).mean())
    B, I, K, H, W = m.shape
    centres = np.zeros((B * I, 2))
    if compact_channel is not None:                                  # compactness_head.py:29-56
        mc = m[:, :, compact_channel].reshape(B * I, H, W)
        cnt = mc.sum(axis=(1, 2), keepdims=True)
        # the reference builds the coordinates in fp32 (torch.arange(..., dtype=torch.float) / mask_H, :38-40)
        y = (np.arange(H, dtype=np.float32) / np.float32(H)).astype(np.float64)[None, :, None]
        xx = (np.arange(W, dtype=np.float32) / np.float32(W)).astype(np.float64)[None, None, :]
        yc = (y * mc).sum(axis=(1, 2), keepdims=True) / cnt
        xc = (xx * mc).sum(axis=(1, 2), keepdims=True) / cnt
        out["compactness"] = float((((y - yc) ** 2 + (xx - xc) ** 2) * mc).mean())
        centres = np.concatenate([yc.reshape(-1, 1), xc.reshape(-1, 1)], 1)
    if target is not None and object_channel is not None:            # rcf_model.py:380-408
        t = _pl_target(target, pl_th)
        d = t - m[:, :, object_channel]
        out["pl"] = float((np.maximum(d, 0) ** 2).mean() * wpos + (np.minimum(d, 0) ** 2).mean() * wneg)
    return m, out, centres


def mask_losses_backward(masks, g_masks=None, g_entropy=None, g_compact=None, g_pl=None, compact_channel=None, target=None,
                         object_channel=None, pl_th=-1.0, wpos=1.0, wneg=1.0, g_sharpen=None, sharpen=None, t_sharpen=0.25):
    """d(loss)/d(logits) for upstream gradients on the masks and on each scalar loss (None = not used)."""
    m = masks
    B, I, K, H, W = m.shape
    npix = B * I * H * W
    g = np.zeros_like(m) if g_masks is None else g_masks.astype(m.dtype).copy()
    if g_entropy is not None:
        lse = np.log(np.exp(m).sum(axis=2, keepdims=True))
        ls = m - lse
        q = np.exp(ls)
        g = g - g_entropy / npix * (ls + m - q * m.sum(axis=2, keepdims=True))
    if g_compact is not None and compact_channel is not None:
        mc = m[:, :, compact_channel].reshape(B * I, H, W)
        cnt = mc.sum(axis=(1, 2), keepdims=True)
        y = (np.arange(H, dtype=np.float32) / np.float32(H)).astype(np.float64)[None, :, None]
        xx = (np.arange(W, dtype=np.float32) / np.float32(W)).astype(np.float64)[None, None, :]
        yc = (y * mc).sum(axis=(1, 2), keepdims=True) / cnt
        xc = (xx * mc).sum(axis=(1, 2), keepdims=True) / cnt
        # the centroid's dependence on m cancels: sum_q m_q (y_q - yc) = 0
        g[:, :, compact_channel] += (g_compact / npix * ((y - yc) ** 2 + (xx - xc) ** 2)).reshape(B, I, H, W)
    if g_pl is not None and target is not None and object_channel is not None:
        d = _pl_target(target, pl_th) - m[:, :, object_channel]
        g[:, :, object_channel] += g_pl / npix * (-2.0) * (wpos * np.maximum(d, 0) + wneg * np.minimum(d, 0))
    if g_sharpen is not None and sharpen == 'kl':
        # target detached; d/d(ls_k) = -t_k / Nel, ls = m - lse(m)  =>  d/dm_j = -(t_j - q_j) / Nel
        q = np.exp(m - np.log(np.exp(m).sum(axis=2, keepdims=True)))
        g = g - g_sharpen / m.size * (_sharpen(m, t_sharpen, 2) - q)
    elif g_sharpen is not None and sharpen == 'object_hinge':
        other = m.copy()
        other[:, :, object_channel] = 0.0
        d = m[:, :, object_channel] - other.max(axis=2)
        g[:, :, object_channel] += g_sharpen / npix * np.where(t_sharpen - np.abs(d) > 0, -np.sign(d), 0.0)
    return m * (g - (m * g).sum(axis=2, keepdims=True))


def mask_prep_forward(logits):
    m, out, _ = mask_losses_forward(logits)
    return m, out["entropy"]


def mask_prep_backward(masks, g_masks, g_entropy):
    return mask_losses_backward(masks, g_masks=g_masks, g_entropy=g_entropy)
