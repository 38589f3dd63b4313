"""Host-side cost of one eager step of the drop-in head at the DAVIS training shape (B=8, 96x96): enqueue time per step and
a cProfile table (what the Python glue, torch.empty and the launches cost when the device time is only ~0.22 ms)."""
import cProfile, pstats, io, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcf_unsupvideoseg_b200 as pkg
dev = torch.device("cuda")
B, K, H, W = 8, 4, 96, 96
head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, mask_layer=K, mask_size=(H, W), clamp_flow_t=20.0, free_residual=True, allow_residual_resize=True).to(dev)
head.return_flows = False
g = torch.Generator(device=dev).manual_seed(0)
masks = torch.softmax(torch.randn(B, 2, K, H, W, device=dev, generator=g) * 2, dim=2).requires_grad_(True)
fw = torch.randn(B, 1, 2, H, W, device=dev, generator=g) * 8
bw = torch.randn(B, 1, 2, H, W, device=dev, generator=g) * 8
r1 = (torch.randn(B, 2 * K, H // 2, W // 2, device=dev, generator=g) * 5).requires_grad_(True)
r2 = (torch.randn(B, 2 * K, H // 2, W // 2, device=dev, generator=g) * 5).requires_grad_(True)
imgs = torch.zeros(B, 2, 3, 8, 8)
params = list(head.parameters())
def step():
    _, l = head(imgs, masks, fw, bw, r1, r2)
    torch.autograd.grad(l["seg"], [masks, r1, r2, *params])
for _ in range(20): step()
torch.cuda.synchronize()
import time
t = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t)/200:.3f} ms/step, wall {1e3*(t2-t)/200:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
