"""Per-kernel device durations of the drop-in head on the REAL timeline (torch.profiler / CUPTI activity records: warm
caches, no serialisation or replay), as opposed to the cold-cache numbers of an ncu launch list.
    python tools/profile_timeline.py --B 8 --H 96 --W 96 --resize [--affine] [--steps 20]
"""
import argparse
import collections
import os
import sys

# per-kernel durations are only exclusive without programmatic dependent launch (a dependent kernel that was scheduled
# early would be charged for the time it spends waiting for its predecessor)
os.environ.setdefault("RCF_PDL", "0")

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcf_unsupvideoseg_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8); ap.add_argument("--K", type=int, default=4)
ap.add_argument("--H", type=int, default=96); ap.add_argument("--W", type=int, default=96)
ap.add_argument("--Cf", type=int, default=64); ap.add_argument("--ks", type=int, default=3)
ap.add_argument("--affine", action="store_true"); ap.add_argument("--resize", action="store_true")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--nprod", type=int, default=0, help="conv_precision of the tcgen05 convs (0 = follow torch settings)")
ap.add_argument("--no-tc", action="store_true", help="second conv through cuDNN (round-1 path)")
ap.add_argument("--cudnn-benchmark", action="store_true", help="torch.backends.cudnn.benchmark = True (algorithm search)")
a = ap.parse_args()
torch.backends.cudnn.benchmark = bool(a.cudnn_benchmark)
dev = torch.device("cuda")
B, K, H, W = a.B, a.K, a.H, a.W
kw = dict(free_residual_with_affine=True) if a.affine else dict(free_residual=True)
torch.manual_seed(1)
head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, mask_layer=K, mask_size=(H, W), clamp_flow_t=20.0,
                                           num_flow_feat_channels=a.Cf, flow_feat_before_agg_kernel_size=a.ks,
                                           allow_residual_resize=a.resize, **kw).to(dev)
head.return_flows = False
head.tensor_core_conv = not a.no_tc
head.conv_precision = a.nprod or None
g = torch.Generator(device=dev).manual_seed(0)
masks = torch.softmax(torch.randn(B, 2, K, H, W, device=dev, generator=g) * 2, dim=2).requires_grad_(True)
fw = torch.randn(B, 1, 2, H, W, device=dev, generator=g) * 8
bw = torch.randn(B, 1, 2, H, W, device=dev, generator=g) * 8
rh, rw = (H // 2, W // 2) if a.resize else (H, W)
r1 = (torch.randn(B, 2 * K, rh, rw, device=dev, generator=g) * 5).requires_grad_(True)
r2 = (torch.randn(B, 2 * K, rh, rw, device=dev, generator=g) * 5).requires_grad_(True)
imgs = torch.zeros(B, 2, 3, 8, 8)
params = list(head.parameters())


def step():
    _, l = head(imgs, masks, fw, bw, r1, r2)
    return torch.autograd.grad(l["seg"], [masks, r1, r2, *params])


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        d = agg.setdefault(ev.name[:96], [0, 0.0])
        d[0] += 1; d[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"# B={B} K={K} {H}x{W} Cf={a.Cf} affine={a.affine} resize={a.resize}: {tot / a.steps:.1f} us of kernel time per step, "
      f"{sum(v[0] for v in agg.values()) / a.steps:.0f} launches per step\n")
print("| kernel | launches/step | us/launch | us/step | share |\n|---|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c / a.steps:.1f} | {t / c:.2f} | {t / a.steps:.1f} | {100 * t / tot:.1f}% |")
