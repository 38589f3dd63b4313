"""Mirror of the reference's utils/warp_utils.py (:8-113) over the sm_100a kernels of csrc/rcf_warp.cu.

Same function names, argument meaning and return shapes: ``mesh_grid``, ``norm_grid``, ``get_corresponding_map``,
``flow_warp``, ``get_occu_mask_bidirection``, ``get_occu_mask_backward``.  The RCF head does not call these
(SURVEY.md 8a row a9); the AMD baseline does (models/amd/pwc_lite.py:199, models/amd/flow_loss.py:73-79).
CUDA tensors only -- there is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib


def mesh_grid(B, H, W):
    """[B,2,H,W] integer pixel grid, channel 0 = x (column), channel 1 = y (row)  (reference :8-14)."""
    xs = torch.arange(0, W).view(1, 1, W).expand(B, H, W)
    ys = torch.arange(0, H).view(1, H, 1).expand(B, H, W)
    return torch.stack([xs, ys], 1)


def norm_grid(v_grid):
    """Scale absolute pixel coordinates [B,2,H,W] to [-1,1] and return [B,H,W,2]  (reference :17-24)."""
    _, _, H, W = v_grid.size()
    out = torch.empty_like(v_grid, dtype=v_grid.dtype if v_grid.is_floating_point() else torch.float32)
    out[:, 0] = 2.0 * v_grid[:, 0] / (W - 1) - 1.0
    out[:, 1] = 2.0 * v_grid[:, 1] / (H - 1) - 1.0
    return out.permute(0, 2, 3, 1)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensors required (the B200 implementation has no CPU fallback)")


class _FlowWarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow, border: bool):
        lib = _lib.load_library()
        _need_cuda(x, "flow_warp")
        xc = x.float().contiguous()
        fc = flow.float().contiguous()
        B, C, H, W = xc.shape
        assert fc.shape == (B, 2, H, W), f"flow shape {tuple(fc.shape)} vs input {tuple(xc.shape)}"
        out = torch.empty_like(xc)
        with _lib.device_guard(xc.device):
            _lib.check(lib.rcf_flow_warp_forward(xc.data_ptr(), fc.data_ptr(), out.data_ptr(), B, C, H, W, int(border),
                                                 _lib.raw_stream(xc.device)), "rcf_flow_warp_forward")
        ctx.save_for_backward(xc, fc)
        ctx.border = border
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        lib = _lib.load_library()
        xc, fc = ctx.saved_tensors
        B, C, H, W = xc.shape
        g = gout.float().contiguous()
        gx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        gf = torch.empty_like(fc) if ctx.needs_input_grad[1] else None
        with _lib.device_guard(xc.device):
            _lib.check(lib.rcf_flow_warp_backward(xc.data_ptr(), fc.data_ptr(), g.data_ptr(),
                                                  gx.data_ptr() if gx is not None else None,
                                                  gf.data_ptr() if gf is not None else None, B, C, H, W, int(ctx.border),
                                                  _lib.raw_stream(xc.device)), "rcf_flow_warp_backward")
        return gx, gf, None


def flow_warp(x, flow12, pad='border', mode='bilinear'):
    """Warp x [B,C,H,W] by flow12 [B,2,H,W] (reference :84-94).  pad: 'border' | 'zeros'; bilinear only."""
    if mode != 'bilinear':
        raise NotImplementedError(f"flow_warp mode={mode!r}: only 'bilinear' is implemented (the only mode the reference uses)")
    if pad not in ('border', 'zeros'):
        raise NotImplementedError(f"flow_warp pad={pad!r}: only 'border' and 'zeros' are implemented")
    out = _FlowWarpFn.apply(x, flow12, pad == 'border')
    return out if out.dtype == x.dtype else out.to(x.dtype)      # grid_sample preserves the input dtype (fp16 under AMP)


def get_corresponding_map(data):
    """Bilinear forward splat of ones: data [B,2,H,W] absolute (x,y) positions -> [B,1,H,W]  (reference :27-81).
    Not differentiable (the reference only thresholds the result)."""
    lib = _lib.load_library()
    _need_cuda(data, "get_corresponding_map")
    d = data.detach().float().contiguous()
    B, _, H, W = d.shape
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=d.device)
    scratch = torch.empty(B * H * W, dtype=torch.int64, device=d.device)
    with _lib.device_guard(d.device):
        _lib.check(lib.rcf_corresponding_map(d.data_ptr(), out.data_ptr(), scratch.data_ptr(), B, H, W,
                                             _lib.raw_stream(d.device)), "rcf_corresponding_map")
    return out.type_as(data)


def get_occu_mask_bidirection(flow12, flow21, scale=0.01, bias=0.5):
    """Forward-backward consistency occlusion mask (reference :97-104)."""
    flow21_warped = flow_warp(flow21, flow12, pad='zeros')
    flow12_diff = flow12 + flow21_warped
    mag = (flow12 * flow12).sum(1, keepdim=True) + (flow21_warped * flow21_warped).sum(1, keepdim=True)
    occ_thresh = scale * mag + bias
    occ = (flow12_diff * flow12_diff).sum(1, keepdim=True) > occ_thresh
    return occ.float()


def get_occu_mask_backward(flow21, th=0.2):
    """Occlusion from the forward-splat density of the backward flow (reference :107-113)."""
    B, _, H, W = flow21.size()
    base_grid = mesh_grid(B, H, W).to(flow21.device).type_as(flow21)
    corr_map = get_corresponding_map(base_grid + flow21)
    occu_mask = corr_map.clamp(min=0., max=1.) < th
    return occu_mask.float()
