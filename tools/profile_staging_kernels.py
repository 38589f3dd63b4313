"""Per-kernel device time and achieved HBM rate of the staging kernels (mask losses, resize, flow staging) on the real
timeline (torch.profiler).   python tools/profile_staging_kernels.py"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

os.environ.setdefault("RCF_PDL", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcf_unsupvideoseg_b200.mask_ops import mask_losses  # noqa: E402
from rcf_unsupvideoseg_b200.resize import resize_bilinear_multi, stage_flow_hwc  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
B, I, K, H, W = 16, 2, 4, 480, 854
logits = (torch.randn(B, I, K, H, W, device=dev, generator=g) * 2).requires_grad_(True)
gm = torch.randn(B, I, K, H, W, device=dev, generator=g)
pl = torch.rand(B, I, H, W, device=dev, generator=g)
gl = torch.ones(3, device=dev)
ra = torch.randn(B, 2 * K, H // 2, W // 2, device=dev, generator=g).requires_grad_(True)
rb = torch.randn(B, 2 * K, H // 2, W // 2, device=dev, generator=g).requires_grad_(True)
go = torch.randn(B, 2 * K, H, W, device=dev, generator=g)
fl = torch.randn(B, H, W, 2, device=dev, generator=g)
npx = B * I * H * W


def step():
    from rcf_unsupvideoseg_b200.mask_ops import _MaskLossesFn
    m, lo, _ = _MaskLossesFn.apply(logits, pl, 0, 1, -1.0, 2.0, 0.5)
    torch.autograd.grad([m, lo], [logits], [gm, gl])
    ya, yb = resize_bilinear_multi([ra, rb], (H, W))
    torch.autograd.grad([ya, yb], [ra, rb], [go, go])
    stage_flow_hwc(fl, (H // 5, W // 5))


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA and ("k_mask" in ev.name or "k_resize" in ev.name or "k_flow_stage" in ev.name):
        d = agg.setdefault(ev.name[:80], [0, 0.0])
        d[0] += 1; d[1] += ev.device_time
bytes_ = {"k_mask_fwd": npx * (4 * K + 4 * K + 4), "k_mask_bwd": npx * (4 * K * 3 + 4),
          "k_resize_fwd": 2 * B * 2 * K * (H * W + H * W // 4) * 4, "k_resize_bwd": 2 * B * 2 * K * (H * W + H * W // 4) * 4,
          "k_flow_stage": B * 2 * (H * W + (H // 5) * (W // 5)) * 4}
print(f"# mask losses (entropy + compactness + PL) on {B}x{I}x{K}x{H}x{W}; resize 2 x {B}x{2*K}x{H//2}x{W//2} -> {H}x{W}; flow staging {B}x{H}x{W}x2 -> {H//5}x{W//5}\n")
print("| kernel | us/launch | algorithmic MB | GB/s | of measured HBM peak (6545 GB/s) |\n|---|---|---|---|---|")
for n, (c, t) in agg.items():
    us = t / c
    key = next((k for k in bytes_ if k in n), None)
    mb = bytes_[key] / 1e6 if key else float("nan")
    gbs = mb / us * 1e3 if key else 0.0       # MB/us = TB/s -> GB/s
    print(f"| `{n}` | {us:.1f} | {mb:.0f} | {gbs:.0f} | {gbs / 6545.3 * 100:.1f}% |" if key else f"| `{n}` | {us:.1f} | - | - | - |")
