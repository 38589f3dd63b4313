"""CPU oracle (numpy fp64) for the reference's utils/warp_utils.py sampling ops -- TEST INFRASTRUCTURE ONLY.

Restates: flow_warp (:84-94, i.e. F.grid_sample bilinear, align_corners=True, padding 'zeros'/'border') with its
gradients w.r.t. the sampled tensor and the flow; get_corresponding_map (:27-81); the two occlusion masks (:97-113).
Pinned to outputs of the unmodified reference functions by tests/golden/warp_*.npz (tests/golden/make_golden_warp.py).
"""
import numpy as np


def _taps(flow, pad):
    B, _, H, W = flow.shape
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    ix = xs[None] + flow[:, 0]
    iy = ys[None] + flow[:, 1]
    sx = np.ones_like(ix); sy = np.ones_like(iy)
    if pad == "border":
        sx = ((ix > 0) & (ix < W - 1)).astype(np.float64)
        sy = ((iy > 0) & (iy < H - 1)).astype(np.float64)
        ix = np.clip(ix, 0, W - 1); iy = np.clip(iy, 0, H - 1)
    x0 = np.floor(ix).astype(np.int64); y0 = np.floor(iy).astype(np.int64)
    return ix - x0, iy - y0, x0, y0, sx, sy


def _gather(x, xi, yi):
    B, C, H, W = x.shape
    ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
    xc = np.clip(xi, 0, W - 1); yc = np.clip(yi, 0, H - 1)
    b = np.arange(B)[:, None, None]
    v = x[b, :, yc, xc]                      # [B,H,W,C]
    return np.moveaxis(v, -1, 1) * ok[:, None], ok


def flow_warp(x, flow, pad="border"):
    x = np.asarray(x, np.float64); flow = np.asarray(flow, np.float64)
    wx, wy, x0, y0, _, _ = _taps(flow, pad)
    v00, _ = _gather(x, x0, y0); v10, _ = _gather(x, x0 + 1, y0)
    v01, _ = _gather(x, x0, y0 + 1); v11, _ = _gather(x, x0 + 1, y0 + 1)
    wx, wy = wx[:, None], wy[:, None]
    return v00 * (1 - wx) * (1 - wy) + v10 * wx * (1 - wy) + v01 * (1 - wx) * wy + v11 * wx * wy


def flow_warp_backward(x, flow, gout, pad="border"):
    x = np.asarray(x, np.float64); flow = np.asarray(flow, np.float64); gout = np.asarray(gout, np.float64)
    B, C, H, W = x.shape
    wx, wy, x0, y0, sx, sy = _taps(flow, pad)
    v00, b00 = _gather(x, x0, y0); v10, b10 = _gather(x, x0 + 1, y0)
    v01, b01 = _gather(x, x0, y0 + 1); v11, b11 = _gather(x, x0 + 1, y0 + 1)
    wxe, wye = wx[:, None], wy[:, None]
    gix = (gout * ((v10 - v00) * (1 - wye) + (v11 - v01) * wye)).sum(1) * sx
    giy = (gout * ((v01 - v00) * (1 - wxe) + (v11 - v10) * wxe)).sum(1) * sy
    gx = np.zeros_like(x)
    b = np.broadcast_to(np.arange(B)[:, None, None], x0.shape)
    for (dx, dy, w, ok) in ((0, 0, (1 - wx) * (1 - wy), b00), (1, 0, wx * (1 - wy), b10),
                            (0, 1, (1 - wx) * wy, b01), (1, 1, wx * wy, b11)):
        xi = np.clip(x0 + dx, 0, W - 1); yi = np.clip(y0 + dy, 0, H - 1)
        for c in range(C):
            np.add.at(gx[:, c], (b, yi, xi), gout[:, c] * w * ok)
    return gx, np.stack([gix, giy], 1)


def get_corresponding_map(data):
    data = np.asarray(data, np.float64)
    B, _, H, W = data.shape
    out = np.zeros((B, H, W))
    x, y = data[:, 0], data[:, 1]
    x1 = np.floor(x); y1 = np.floor(y)
    wx, wy = x - x1, y - y1
    b = np.broadcast_to(np.arange(B)[:, None, None], x.shape)
    for (dx, dy, w) in ((0, 0, (1 - wx) * (1 - wy)), (1, 0, wx * (1 - wy)), (0, 1, (1 - wx) * wy), (1, 1, wx * wy)):
        xi = (x1 + dx).astype(np.int64); yi = (y1 + dy).astype(np.int64)
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        np.add.at(out, (b[ok], yi[ok], xi[ok]), w[ok])
    return out[:, None]


def get_occu_mask_backward(flow21, th=0.2):
    B, _, H, W = flow21.shape
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    grid = np.stack([xs, ys])[None] + np.asarray(flow21, np.float64)
    return (np.clip(get_corresponding_map(grid), 0, 1) < th).astype(np.float64)


def get_occu_mask_bidirection(flow12, flow21, scale=0.01, bias=0.5):
    flow12 = np.asarray(flow12, np.float64)
    w = flow_warp(flow21, flow12, pad="zeros")
    diff = flow12 + w
    mag = (flow12 ** 2).sum(1, keepdims=True) + (w ** 2).sum(1, keepdims=True)
    return ((diff ** 2).sum(1, keepdims=True) > scale * mag + bias).astype(np.float64)
