"""BASELINE.json configs C2/C4/C5 (+ affine) through the loss core (theta supplied), CUDA-graph replay, one table.
    python tools/sweep.py > gpurun_out/sweep.md
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcf_unsupvideoseg_b200 as pkg  # noqa: E402


def run(B, K, H, W, D, robust=False, steps=30):
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    masks = torch.softmax(torch.randn(B, 2, K, H, W, device=dev, generator=g) * 2, dim=2).requires_grad_(True)
    flows = [torch.randn(B, 2, H, W, device=dev, generator=g) * 8 for _ in range(2)]
    resids = [(torch.randn(B, 2 * K, H, W, device=dev, generator=g) * 5).requires_grad_(True) for _ in range(2)]
    thetas = [torch.randn(B, 2, K, device=dev, generator=g).requires_grad_(True) for _ in range(2)]
    spec = pkg.LossSpec(K=K, H=H, W=W, D=D, Cf=0, clamp_t=20.0, robust=robust)
    gl = torch.ones(2, device=dev)

    def step():
        loss, _ = pkg.rcf_motion_loss(spec, masks, flows, resids, thetas=thetas)
        return loss, torch.autograd.grad(loss, [masks, *resids, *thetas], grad_outputs=gl)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = step()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        gr.replay()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    del out, gr
    torch.cuda.empty_cache()
    return ms


def main():
    peak = 6545.3
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    rows = [("C2", 16, 4, 480, 854, 0, False), ("C2 affine (STv2/FBMS flags)", 16, 4, 480, 854, 2, False),
            ("C2 robust loss", 16, 4, 480, 854, 0, True),
            ("C3 per-GPU shard B=32 (2 GPUs)", 32, 4, 480, 854, 0, False), ("C3 per-GPU shard B=8 (8 GPUs)", 8, 4, 480, 854, 0, False),
            ("C4 K=2", 32, 2, 480, 854, 0, False), ("C4 K=4", 32, 4, 480, 854, 0, False), ("C4 K=8", 32, 8, 480, 854, 0, False),
            ("C4 K=8 affine", 32, 8, 480, 854, 2, False), ("FBMS K=3 affine", 16, 3, 480, 854, 2, False),
            ("C5 240x427", 16, 4, 240, 427, 0, False), ("C5 480x854", 16, 4, 480, 854, 0, False),
            ("C5 1080x1920", 16, 4, 1080, 1920, 0, False), ("quadratic fit (never enabled in configs)", 16, 4, 480, 854, 5, False),
            ("DAVIS training shape 96x96 B=8", 8, 4, 96, 96, 0, False), ("STv2 training shape 48x48 B=8 affine", 8, 4, 48, 48, 2, False)]
    print(f"| config | B | K | HxW | mode | ms/step (fwd+bwd) | samples/s | algorithmic GB/s | of measured HBM peak ({peak:.0f} GB/s) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, B, K, H, W, D, robust in rows:
        ms = run(B, K, H, W, D, robust)
        alg = B * 2 * H * W * (36 * K + 16)
        gbs = alg / ms / 1e6
        mode = {0: "free", 2: "affine", 5: "quadratic"}[D] + ("+robust" if robust else "")
        print(f"| {name} | {B} | {K} | {H}x{W} | {mode} | {ms:.4f} | {B / ms * 1e3:.0f} | {gbs:.0f} | {100 * gbs / peak:.1f}% |", flush=True)


if __name__ == "__main__":
    main()
