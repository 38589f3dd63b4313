"""Time the hand-written conv stem (fwd / bwd) alone.  python tools/profile_stem.py --B 4 --H 480 --W 854"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcf_unsupvideoseg_b200.stem import flow_stem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=2); ap.add_argument("--H", type=int, default=480); ap.add_argument("--W", type=int, default=854)
ap.add_argument("--Cf", type=int, default=64); ap.add_argument("--ks", type=int, default=3)
a = ap.parse_args()
flows = [torch.randn(a.B, 2, a.H, a.W, device="cuda") * 8 for _ in range(2)]
conv = torch.nn.Conv2d(2, a.Cf, a.ks, padding=(a.ks - 1) // 2).cuda()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / n
fwd = t(lambda: flow_stem(flows, conv.weight, conv.bias, 20.0, 0.1))
act = flow_stem(flows, conv.weight, conv.bias, 20.0, 0.1)
g = torch.randn_like(act)
both = t(lambda: torch.autograd.grad(flow_stem(flows, conv.weight, conv.bias, 20.0, 0.1), [conv.weight, conv.bias], g))
px = 2 * a.B * a.H * a.W
print(f"stem fwd {fwd*1e3:.1f} us ({px*(8+4*a.Cf)/fwd/1e6:.0f} GB/s), bwd {(both-fwd)*1e3:.1f} us ({px*(8+8*a.Cf)/(both-fwd)/1e6:.0f} GB/s)")
