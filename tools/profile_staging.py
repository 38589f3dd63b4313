"""Caller-side staging ops against the ATen op sequences of the reference, CUDA-event timed.
    python tools/profile_staging.py
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcf_unsupvideoseg_b200.mask_ops import mask_losses  # noqa: E402
from rcf_unsupvideoseg_b200.resize import resize_bilinear_multi  # noqa: E402


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:                                   # device time without the Python launch overhead
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            fn()
        run = g.replay
    except Exception:                      # noqa: BLE001
        run = fn
    for _ in range(3):
        run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        run()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def reference_mask_ops(logits, w_mask, pl, cc, oc):
    """models/rcf_model.py:433-434, :376-378, :380-393 and compactness_head.py:29-56 as ATen ops (+ backward)."""
    m = F.softmax(logits, dim=2)
    lm = F.log_softmax(m, dim=2)
    ent = -(m * lm).sum(dim=2).mean()
    mc = m.flatten(0, 1)[:, cc]
    H, W = mc.shape[-2:]
    cnt = mc.sum(dim=(1, 2), keepdim=True)
    y = torch.arange(H, dtype=torch.float, device=m.device)[None, :, None] / H
    x = torch.arange(W, dtype=torch.float, device=m.device)[None, None, :] / W
    yc = (y * mc).sum(dim=(1, 2), keepdim=True) / cnt
    xc = (x * mc).sum(dim=(1, 2), keepdim=True) / cnt
    comp = (((y - yc) ** 2 + (x - xc) ** 2) * mc).mean()
    d = pl - m[:, :, oc]
    plo = (torch.clamp(d, min=0) ** 2).mean() * 2.0 + (torch.clamp(d, max=0) ** 2).mean() * 0.5
    return torch.autograd.grad((m * w_mask).sum() + 0.05 * ent + comp + 2.0 * plo, logits)


def ours_mask_ops(logits, w_mask, pl, cc, oc):
    m, lo = mask_losses(logits, compact_channel=cc, pl_masks=pl, object_channel=oc, pl_pos_weight=2.0, pl_neg_weight=0.5)
    return torch.autograd.grad((m * w_mask).sum() + 0.05 * lo["entropy"] + lo["compactness"] + 2.0 * lo["pl"], logits)


print("| op | shape | ATen (reference op sequence) | ours | speed-up |\n|---|---|---|---|---|")
for shape in ((8, 2, 4, 96, 96), (8, 2, 4, 48, 48), (16, 2, 4, 480, 854)):
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = (torch.randn(*shape, device="cuda", generator=g) * 2).requires_grad_(True)
    w_mask = torch.randn(*shape, device="cuda", generator=g)
    pl = torch.rand(shape[0], shape[1], shape[3], shape[4], device="cuda", generator=g)
    t_ref = timeit(lambda: reference_mask_ops(logits, w_mask, pl, 0, 1))
    t_our = timeit(lambda: ours_mask_ops(logits, w_mask, pl, 0, 1))
    print(f"| softmax + entropy + compactness + PL loss, fwd+bwd | {shape} | {t_ref:.1f} us | {t_our:.1f} us | {t_ref / t_our:.1f} x |")
for (B, C, h, w, H, W) in ((8, 8, 48, 48, 96, 96), (16, 8, 240, 427, 480, 854)):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(B, C, h, w, device="cuda", generator=g).requires_grad_(True)
    b = torch.randn(B, C, h, w, device="cuda", generator=g).requires_grad_(True)
    go = torch.randn(B, C, H, W, device="cuda", generator=g)

    def ref():
        ya = F.interpolate(a, (H, W), mode="bilinear"); yb = F.interpolate(b, (H, W), mode="bilinear")
        return torch.autograd.grad([ya, yb], [a, b], [go, go])

    def ours():
        ya, yb = resize_bilinear_multi([a, b], (H, W))
        return torch.autograd.grad([ya, yb], [a, b], [go, go])

    t_ref, t_our = timeit(ref), timeit(ours)
    print(f"| bilinear resize of both residual maps, fwd+bwd | 2x{(B, C, h, w)} -> {(H, W)} | {t_ref:.1f} us | {t_our:.1f} us | {t_ref / t_our:.1f} x |")
