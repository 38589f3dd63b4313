"""Run a few device-resident loss-core steps (the bench.py `value` path) -- target for ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:k_bwd -s 2 -c 2 -o gpurun_out/prof \
        python tools/profile_step.py --steps 3
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcf_unsupvideoseg_b200 as pkg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--K", type=int, default=4)
    ap.add_argument("--H", type=int, default=480)
    ap.add_argument("--W", type=int, default=854)
    ap.add_argument("--D", type=int, default=0)
    ap.add_argument("--robust", action="store_true")
    ap.add_argument("--Cf", type=int, default=0, help=">0: Scope H, pool a random [B,Cf,H,W] feature map + segment MLP")
    ap.add_argument("--nhwc", action="store_true", help="feature map in channels-last layout")
    ap.add_argument("--slope", type=float, default=1.0, help="fused LeakyReLU slope of the feature map (1 = none)")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--time", action="store_true", help="print CUDA-event time per step and per kernel")
    ap.add_argument("--graph", action="store_true", help="also time the step replayed from a CUDA graph")
    ap.add_argument("--opt", type=str, default="", help="comma list of option=value for rcf_debug_set_option, e.g. 3=0,5=0 (no L2 hints, no programmatic dependent launch)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    B, K, H, W = a.B, a.K, a.H, a.W
    masks = torch.softmax(torch.randn(B, 2, K, H, W, device=dev, generator=g) * 2, dim=2).requires_grad_(True)
    flows = [torch.randn(B, 2, H, W, device=dev, generator=g) * 8 for _ in range(2)]
    resids = [(torch.randn(B, 2 * K, H, W, device=dev, generator=g) * 5).requires_grad_(True) for _ in range(2)]
    thetas = [torch.randn(B, 2, K, device=dev, generator=g).requires_grad_(True) for _ in range(2)]
    spec = pkg.LossSpec(K=K, H=H, W=W, D=a.D, Cf=a.Cf, clamp_t=20.0, robust=a.robust, feat_lrelu_slope=a.slope)
    gl = torch.ones(2, device=dev)
    lib = pkg.load_library()
    for kv in filter(None, a.opt.split(",")):
        o, v = kv.split("=")
        assert lib.rcf_debug_set_option(int(o), int(v)) == 0
    Cf = a.Cf
    if Cf > 0:
        feat = torch.randn(2 * B, Cf, H, W, device=dev, generator=g)
        if a.nhwc:
            feat = feat.contiguous(memory_format=torch.channels_last)
        feat = feat.view(2, B, Cf, H, W).requires_grad_(True)
        mlp = [(torch.randn(Cf, Cf, 1, device=dev, generator=g) / Cf ** 0.5).requires_grad_(True),
               torch.zeros(Cf, device=dev).requires_grad_(True),
               (torch.randn(2, Cf, 1, device=dev, generator=g) / Cf ** 0.5).requires_grad_(True),
               torch.zeros(2, device=dev).requires_grad_(True)]

    def step():
        if Cf > 0:
            loss, _ = pkg.rcf_motion_loss(spec, masks, flows, resids, feats=feat, mlp=mlp)
            return loss, torch.autograd.grad(loss, [masks, *resids, feat, *mlp], grad_outputs=gl)
        loss, _ = pkg.rcf_motion_loss(spec, masks, flows, resids, thetas=thetas)
        return loss, torch.autograd.grad(loss, [masks, *resids, *thetas], grad_outputs=gl)

    if not a.time:
        for _ in range(a.steps):
            loss, _ = step()
        torch.cuda.synchronize()
        print("loss", loss.tolist())
        return
    P = H * W
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    s.record()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    cpu_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.steps
    alg = B * 2 * P * (36 * K + 16 + 12 * Cf + (16 * K if Cf else 0))
    print(f"step {ms:.4f} ms  {B / ms * 1e3:.0f} samples/s  algorithmic {alg / ms / 1e6:.0f} GB/s  (cpu enqueue {cpu_ms:.3f} ms/step)")
    if a.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss_g, grads_g = step()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        s.record()
        for _ in range(a.steps):
            g.replay()
        e.record(); torch.cuda.synchronize()
        msg = s.elapsed_time(e) / a.steps
        print(f"graph step {msg:.4f} ms  {B / msg * 1e3:.0f} samples/s  algorithmic {alg / msg / 1e6:.0f} GB/s  loss {loss_g.tolist()}")
    names = {1: ("k_moments", 4 * K + (8 if a.D else 0)), 2: ("k_loss", 12 * K + 8), 3: ("k_bwd", 24 * K + 8)}
    if Cf > 0:
        names[4] = ("k_pool", 4 * Cf + 4 * K)
        names[5] = ("k_pool_bwd", 8 * Cf + 12 * K)
    for which, (name, bpp) in names.items():
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        for x, y in evs:
            x.record(); y.record()
        for x, y in evs:
            lib.rcf_debug_time_kernel(which, x.cuda_event, y.cuda_event)
            step()
        lib.rcf_debug_time_kernel(0, None, None)
        torch.cuda.synchronize()
        t = sorted(x.elapsed_time(y) for x, y in evs)
        mean = sum(t) / len(t)
        print(f"  {name:10s} mean {mean * 1e3:8.1f} us  min {t[0] * 1e3:8.1f} us  {B * 2 * P * bpp / mean / 1e6:7.0f} GB/s (algorithmic {bpp} B/px)")


if __name__ == "__main__":
    main()
