"""Debug: which bench.py ingredient breaks CUDA-graph capture / slows the eager loop."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rcf_unsupvideoseg_b200 as pkg

variant = sys.argv[1]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = pkg.load_library()
B, K, H, W = 16, 4, 480, 854
from oracle.torch_port import synthetic_inputs
if "gpu_inputs" in variant:
    g = torch.Generator(device=dev).manual_seed(0)
    masks = torch.softmax(torch.randn(B, 2, K, H, W, device=dev, generator=g) * 2, dim=2).requires_grad_(True)
    fw = torch.randn(B, 1, 2, H, W, device=dev, generator=g) * 8
    bw = torch.randn(B, 1, 2, H, W, device=dev, generator=g) * 8
    rfw = (torch.randn(B, 2 * K, H, W, device=dev, generator=g) * 5).requires_grad_(True)
    rbw = (torch.randn(B, 2 * K, H, W, device=dev, generator=g) * 5).requires_grad_(True)
else:
    masks_h, fw_h, bw_h, rfw_h, rbw_h = synthetic_inputs(B, K, H, W, seed=0)
    masks = masks_h.to(dev).requires_grad_(True)
    fw, bw = fw_h.to(dev), bw_h.to(dev)
    rfw = rfw_h.to(dev).requires_grad_(True)
    rbw = rbw_h.to(dev).requires_grad_(True)
thetas = [torch.randn(B, 2, K).to(dev).requires_grad_(True) for _ in range(2)]
inv_n = 1.0 / (B * 2 * H * W) if "inv_n" in variant else 0.0
spec = pkg.LossSpec(K=K, H=H, W=W, D=0, Cf=0, clamp_t=20.0, inv_n=inv_n)
flows = [fw[:, 0], bw[:, 0]]
gl = torch.ones(2, device=dev)
inputs = [masks, rfw, rbw, *thetas]

def step():
    loss, _ = pkg.rcf_motion_loss(spec, masks, flows, [rfw, rbw], thetas=thetas)
    return loss, torch.autograd.grad(loss, inputs, grad_outputs=gl)

if "nvml" in variant:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    t0 = time.perf_counter(); pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); t1 = time.perf_counter()
    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h); t2 = time.perf_counter()
    print(f"nvml clock query {1e3*(t1-t0):.2f} ms, reasons query {1e3*(t2-t1):.2f} ms")
for _ in range(5):
    step()
torch.cuda.synchronize()
steps = 30
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for a, b in evs:
    a.record(); b.record()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); t0 = time.perf_counter()
for i in range(steps):
    if "hook" in variant:
        lib.rcf_debug_time_kernel(3, evs[i][0].cuda_event, evs[i][1].cuda_event)
    if "nvmlloop" in variant and i % 10 == 0:
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    loss, grads = step()
cpu = (time.perf_counter() - t0) * 1e3 / steps
e.record(); lib.rcf_debug_time_kernel(0, None, None); torch.cuda.synchronize()
print(f"[{variant}] eager {s.elapsed_time(e)/steps:.4f} ms/step, cpu enqueue {cpu:.3f} ms/step")
if "dropref" in variant:
    del loss, grads
try:
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        lg, gg = step()
    g_.replay(); torch.cuda.synchronize()
    s.record()
    for _ in range(steps):
        g_.replay()
    e.record(); torch.cuda.synchronize()
    print(f"[{variant}] graph OK {s.elapsed_time(e)/steps:.4f} ms/step")
except Exception as ex:
    print(f"[{variant}] graph FAILED: {str(ex)[:200]}")
    traceback.print_exc(limit=6)
