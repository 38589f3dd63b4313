"""Run the drop-in head (full feature branch) fwd+bwd for a few steps -- target for ncu launch lists.
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l.csv \
        python tools/profile_head.py --B 8 --H 96 --W 96 --steps 3
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcf_unsupvideoseg_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--H", type=int, default=96)
    ap.add_argument("--W", type=int, default=96)
    ap.add_argument("--Cf", type=int, default=64)
    ap.add_argument("--ks", type=int, default=3)
    ap.add_argument("--affine", action="store_true")
    ap.add_argument("--resize", action="store_true", help="residual at half resolution + allow_residual_resize")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--nprod", type=int, default=0, help="conv_precision of the tcgen05 convs (0 = follow torch settings)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    B, K, H, W = a.B, a.K, a.H, a.W
    kw = dict(free_residual_with_affine=True) if a.affine else dict(free_residual=True)
    torch.manual_seed(1)
    head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, mask_layer=K, mask_size=(H, W),
                                               clamp_flow_t=20.0, num_flow_feat_channels=a.Cf,
                                               flow_feat_before_agg_kernel_size=a.ks,
                                               allow_residual_resize=a.resize, **kw).to(dev)
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

    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    if a.time:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(50):
            step()
        e.record(); torch.cuda.synchronize()
        print(f"eager head step {s.elapsed_time(e) / 50:.4f} ms")


if __name__ == "__main__":
    main()
