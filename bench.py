#!/usr/bin/env python
"""bench.py -- RCF motion loss fwd+bwd throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (config C2 of BASELINE.json, per GPU): B=16 samples (frame pairs), K=4 soft masks, 480x854,
stage-1 DAVIS flags (free_residual, L1, clamp_flow_t=20, s=tau=10; configs/rcf/rcf_stage1.yaml:99-111),
synthetic inputs of SURVEY.md 8(d).  Weak scaling: every rank owns its own B=16 shard; the mean's
normaliser is the global element count, so the per-rank losses SUM to the global loss with one NCCL
all-reduce of 2 floats per step (issued asynchronously, off the critical path).

  value  : device-resident loss core (theta supplied = SURVEY 8(d) "Scope L"): rcf_forward + rcf_backward
           through the C ABI, samples/s over all ranks, CUDA-event timed, max over ranks.
  e2e    : the drop-in nn.Module call (FlowAggregationHeadWithResidual with the loss-core proxy feature
           branch Cf=2,k=1 of SURVEY 8(d)) fed from PINNED HOST buffers every step: H2D of masks/flows/
           residuals, fwd+bwd, D2H of the loss scalars and of the mask/residual gradients.
  roofline: dominant kernel k_bwd (streaming backward): algorithmic bytes per launch / mean CUDA-event
           duration of that kernel inside the timed region (library timing hook), vs MEASURED_PEAKS.json.
  cpu_baseline: oracle/torch_port.py (op-for-op PyTorch CPU port of the reference head, pinned to the
           reference's outputs by tests) on the host cores, bounded sample (B=2), rank 0, N=1 only.
--impl reference times that CPU port alone on the same config (the reference itself is Python under
/root/reference, which does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RCF loss fwd+bwd frames/sec and % HBM roofline at 1/2/4/8 B200 vs host-CPU ref"
UNIT = "samples/s"
B_PER_GPU, K, H, W = 16, 4, 480, 854
HEAD_KW = dict(mask_layer=K, mask_size=(H, W), free_residual=True, clamp_flow_t=20.0,
               num_flow_feat_channels=2, flow_feat_before_agg_kernel_size=1)


def workload_config(n_gpus):
    return {
        "workload": f"C2: RCF loss fwd+bwd, B={B_PER_GPU}/GPU, K={K}, {H}x{W}, free_residual+L1+clamp20 (stage-1 DAVIS flags)",
        "value_path": "loss core, theta supplied (SURVEY 8d Scope L), C ABI rcf_forward+rcf_backward, device-resident",
        "e2e_path": "drop-in head, proxy feature branch Cf=2 k=1, pinned host buffers in, loss+grads out",
        "global_batch": B_PER_GPU * n_gpus, "parallelism": f"batch-sharded x{n_gpus}",
        "l2_policy": "inputs (735 MB/GPU) larger than L2 (126 MB); no flush needed",
        "algorithmic_bytes_per_sample": 2 * H * W * (36 * K + 16),
    }


def read_traffic(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    try:
        with open(p) as f:
            for name, rec in json.load(f).items():
                if kernel_prefix in name:
                    return rec["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    return None


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons through NVML (what nvidia-smi reads).  Sampling is done from the main
    thread AFTER a timed region's work has been enqueued and WHILE the GPU is still executing it
    (`poll_until(end_event)`): NVML queries from a second thread during the enqueue loop were measured to
    stall kernel launches (driver lock) and to distort a CPU-sensitive loop."""

    NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.sample(record=False)      # first NVML queries are slow: pay for them outside the timed regions
        except Exception:  # noqa: BLE001
            self.nv = None

    def sample(self, record=True):
        if self.nv is None:
            return
        try:
            mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            try:
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:  # noqa: BLE001
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            if record:
                self.samples.append(mhz)
                for bit, name in self.NAMES.items():
                    if mask & bit:
                        self.reasons.add(name)
        except Exception:  # noqa: BLE001
            pass

    def poll_until(self, end_event, period_s=0.01):
        """Sample while the GPU drains the enqueued region; returns when `end_event` has completed."""
        while True:
            self.sample()
            if end_event.query():
                return
            time.sleep(period_s)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "how": "NVML, sampled under load between enqueue and completion of each timed region"}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------------
def cpu_port_time(steps, warmup, budget_s=150.0):
    """Times oracle/torch_port.py (the reference's op sequence in PyTorch CPU) on a bounded sample."""
    import torch

    from oracle.torch_port import PortedHead, synthetic_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = 2
    torch.manual_seed(1)
    head = PortedHead(**HEAD_KW)
    masks, fw, bw, rfw, rbw = synthetic_inputs(Bs, K, H, W, seed=0)
    masks.requires_grad_(True); rfw.requires_grad_(True); rbw.requires_grad_(True)
    imgs = torch.zeros(Bs, 2, 3, 8, 8)

    def step():
        for t in (masks, rfw, rbw):
            t.grad = None
        _, loss = head(imgs, masks, fw, bw, rfw, rbw)
        loss["seg"].backward()
        return float(loss["seg"].detach())

    t0 = time.perf_counter(); step(); first = time.perf_counter() - t0
    steps = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    for _ in range(max(0, warmup - 1)):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(); times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return {"value": Bs / mean, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"B={Bs} of the C2 workload ({K=}, {H}x{W}, proxy head Cf=2 k=1), {len(times)} timed steps, "
                      f"mean {mean * 1e3:.1f} ms/step, best {min(times) * 1e3:.1f} ms; torch {torch.__version__} CPU, "
                      f"{cores} threads"}, mean, steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, mean, steps = cpu_port_time(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import rcf_unsupvideoseg_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout clean for the one JSON line: NCCL's banner / debug output goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/rcf_bench_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)
    lib = pkg.load_library(build_if_missing=False)

    B = B_PER_GPU
    P = H * W
    inv_n = 1.0 / (B * world * 2 * P)
    from oracle.torch_port import synthetic_inputs   # input generator only (torch CPU generator)
    masks_h, fw_h, bw_h, rfw_h, rbw_h = synthetic_inputs(B, K, H, W, seed=rank)
    masks = masks_h.to(dev).requires_grad_(True)
    fw, bw = fw_h.to(dev), bw_h.to(dev)
    rfw = rfw_h.to(dev).requires_grad_(True)
    rbw = rbw_h.to(dev).requires_grad_(True)
    g = torch.Generator().manual_seed(100 + rank)
    thetas = [torch.randn(B, 2, K, generator=g).to(dev).requires_grad_(True) for _ in range(2)]
    spec = pkg.LossSpec(K=K, H=H, W=W, D=0, Cf=0, clamp_t=20.0, inv_n=inv_n)
    flows = [fw[:, 0], bw[:, 0]]
    gl = torch.ones(2, device=dev)
    inputs = [masks, rfw, rbw, *thetas]
    pending = []

    def step():
        loss, _ = pkg.rcf_motion_loss(spec, masks, flows, [rfw, rbw], thetas=thetas)
        grads = torch.autograd.grad(loss, inputs, grad_outputs=gl)
        if world > 1:
            lr = loss.detach().clone()
            pending.append((dist.all_reduce(lr, async_op=True), lr))
        return loss, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Timed region A (eager): K steps with the library's timing hook around k_bwd -> roofline of the dominant kernel.
    # Timed region B (CUDA graph): the same step captured once and replayed K times -> `value`
    # (falls back to region A's time if capture is not possible).
    ev_pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev_pairs:      # materialise the handles
        a.record(); b.record()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local_rank)
    barrier()
    start.record()
    t0 = time.perf_counter()
    ev_handles = [(a.cuda_event, b.cuda_event) for a, b in ev_pairs]
    for i in range(args.steps):
        lib.rcf_debug_time_kernel(3, ev_handles[i][0], ev_handles[i][1])
        loss, grads = step()
    cpu_enqueue_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    end.record()
    lib.rcf_debug_time_kernel(0, None, None)
    clocks.poll_until(end)
    barrier()
    for w, _ in pending:
        w.wait()
    pending.clear()
    eager_ms = start.elapsed_time(end) / args.steps
    loss_val = [float(x) for x in loss.detach().cpu()]
    # Drop every reference to the eager steps before capturing: a live autograd graph pins the leaves'
    # AccumulateGrad nodes to the default stream, and the engine's stream hand-off then invalidates the capture.
    del loss, grads
    import gc
    gc.collect()
    torch.cuda.synchronize()

    graph_ms = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        for w, _ in pending:
            w.wait()
        pending.clear()
        gc.collect()
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            loss_g, _ = pkg.rcf_motion_loss(spec, masks, flows, [rfw, rbw], thetas=thetas)
            grads_g = torch.autograd.grad(loss_g, inputs, grad_outputs=gl)
        for _ in range(3):
            g_.replay()
        barrier()
        gs_, ge_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gs_.record()
        for _ in range(args.steps):
            g_.replay()
            if world > 1:      # the per-step loss all-reduce stays outside the graph, asynchronous
                lr = loss_g.detach().clone()
                pending.append((dist.all_reduce(lr, async_op=True), lr))
        ge_.record()
        clocks.poll_until(ge_)
        barrier()
        for w, _ in pending:
            w.wait()
        pending.clear()
        graph_ms = gs_.elapsed_time(ge_) / args.steps
        lfin = loss_g.detach().clone()
        if world > 1:
            dist.all_reduce(lfin)
        loss_val = [float(x) for x in lfin.cpu()]
    except Exception as ex:  # noqa: BLE001
        print(f"[bench] CUDA graph capture unavailable ({type(ex).__name__}: {str(ex)[:300]}); using eager timing",
              file=sys.stderr)
        torch.cuda.synchronize()
    best_ms = eager_ms if graph_ms is None else min(graph_ms, eager_ms)
    t = torch.tensor([best_ms, eager_ms, graph_ms if graph_ms is not None else -1.0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, eager_ms, graph_ms = float(t[0]), float(t[1]), float(t[2])
    value = B * world / (ms_step * 1e-3)
    kb_ms = sorted(a.elapsed_time(b) for a, b in ev_pairs)
    kb_mean = sum(kb_ms) / len(kb_ms)

    # e2e through the drop-in module, pinned host buffers in / loss + grads out
    torch.manual_seed(1)
    head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, **HEAD_KW).to(dev)
    head.return_flows = False
    head.loss_inv_n = inv_n
    pin = [t.pin_memory() for t in (masks_h, fw_h, bw_h, rfw_h, rbw_h)]
    out_pin = [torch.empty(2, dtype=torch.float32).pin_memory(), torch.empty_like(masks_h).pin_memory(),
               torch.empty_like(rfw_h).pin_memory(), torch.empty_like(rbw_h).pin_memory()]
    imgs = torch.zeros(B, 2, 3, 8, 8)
    h2d = sum(t.numel() * 4 for t in pin)
    d2h = sum(t.numel() * 4 for t in out_pin)
    # Three streams, double-buffered device inputs: the H2D copy of step i+1 and the D2H copy of step i-1 overlap
    # the compute of step i (PCIe is full duplex; both copy engines busy).  Every step still moves all of its
    # inputs host->device and all of its results device->host inside the timed region.
    s_in, s_out, s_cmp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
    NBUF = 2
    d_in = [[torch.empty_like(t, device=dev) for t in pin] for _ in range(NBUF)]
    ev_in = [torch.cuda.Event() for _ in range(NBUF)]
    ev_free = [torch.cuda.Event() for _ in range(NBUF)]      # compute finished reading buffer b
    ev_out_done = torch.cuda.Event()
    keep = []

    def e2e_step(i):
        bsel = i % NBUF
        with torch.cuda.stream(s_in):
            if i >= NBUF:
                s_in.wait_event(ev_free[bsel])
            for d, src in zip(d_in[bsel], pin):
                d.copy_(src, non_blocking=True)
            ev_in[bsel].record(s_in)
        s_cmp.wait_event(ev_in[bsel])
        m = d_in[bsel][0].requires_grad_(True)
        r1 = d_in[bsel][3].requires_grad_(True)
        r2 = d_in[bsel][4].requires_grad_(True)
        _, fl = head(imgs, m, d_in[bsel][1], d_in[bsel][2], r1, r2)
        lvec = torch.stack([fl["seg_fw"], fl["seg_bw"]]).detach()
        gm, g1, g2 = torch.autograd.grad(fl["seg"], [m, r1, r2])
        ev_free[bsel].record(s_cmp)
        ev_cmp = torch.cuda.Event()
        ev_cmp.record(s_cmp)
        for t_ in (d_in[bsel][0], d_in[bsel][3], d_in[bsel][4]):
            t_.requires_grad_(False)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp)
            for dst, src in zip(out_pin, (lvec, gm, g1, g2)):
                dst.copy_(src, non_blocking=True)
                src.record_stream(s_out)
            ev_out_done.record(s_out)
        keep.append((lvec, gm, g1, g2))
        if len(keep) > 2:
            keep.pop(0)

    e2e_steps = max(3, min(args.steps, 20))
    for i in range(4):
        e2e_step(i)
    s_cmp.wait_event(ev_out_done)
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for i in range(e2e_steps):
        e2e_step(i)
    s_cmp.wait_stream(s_out)
    s_cmp.wait_stream(s_in)
    e2.record()
    clocks.poll_until(e2)
    barrier()
    keep.clear()
    t2 = torch.tensor([s2.elapsed_time(e2)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2) / e2e_steps
    e2e_val = B * world / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = read_peaks()
        bytes_sample = 2 * P * (36 * K + 16)
        kb_bytes = B * 2 * P * ((4 * K + 8 + 8 * K) + (4 * K + 8 * K))       # k_bwd: re-read inputs + write dM, dR
        achieved = kb_bytes / (kb_mean * 1e-3) / 1e9
        step_gbs = B * bytes_sample / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "pcie_gbs_each_way": [h2d / e2e_ms / 1e6, d2h / e2e_ms / 1e6],
                    "note": "PCIe-bound; copies of neighbouring steps overlap compute on 3 streams"},
            "gpu_launches": 5 * args.steps,   # k_loss, k_finalize, k_loss_sum | k_segment_bwd, k_bwd (single-pass forward)
            "roofline": {"bound": "hbm", "kernel": "k_bwd<K=4,D=0,PX=4> (streaming backward)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": read_traffic("k_bwd<4, 0, 4>"),
                         "peak_source": peak_src, "kernel_ms": kb_mean, "kernel_ms_min": kb_ms[0],
                         "algorithmic_bytes_per_launch": kb_bytes},
            "step_roofline": {"algorithmic_gbs": step_gbs, "frac": step_gbs / peak,
                              "frame_directions_per_s": 2 * value},
            "timing": {"eager_ms_per_step": eager_ms, "cuda_graph_ms_per_step": (graph_ms if graph_ms > 0 else None),
                       "cpu_enqueue_ms_per_step": cpu_enqueue_ms,
                       "value_from": "cuda_graph" if (graph_ms > 0 and graph_ms <= eager_ms) else "eager"},
            "loss": loss_val,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _, _ = cpu_port_time(20, 2, budget_s=25.0)
            if not args.no_extras:
                line["extras"] = extras(pkg, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _time_cuda(fn, steps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def extras(pkg, dev):
    """Add-on measurements reported beside the headline (SURVEY.md 8(d) 'report separately'):
    the same workloads through (a) the drop-in head and (b) the op-for-op PyTorch port of the reference
    running EAGER ON THE SAME GPU (the GPU baseline the reference user has today)."""
    import torch

    from oracle.torch_port import PortedHead, synthetic_inputs
    out = {}
    cases = {
        "c2_proxy_head_B16_480x854": (16, 4, 480, 854, dict(free_residual=True, clamp_flow_t=20.0, num_flow_feat_channels=2,
                                                             flow_feat_before_agg_kernel_size=1), 10),
        "c2_affine_proxy_head_B16_480x854": (16, 4, 480, 854, dict(free_residual_with_affine=True, clamp_flow_t=20.0,
                                                                   num_flow_feat_channels=2,
                                                                   flow_feat_before_agg_kernel_size=1), 10),
        "c1_full_head_B2_480x854": (2, 4, 480, 854, dict(free_residual=True, clamp_flow_t=20.0), 5),
        "davis_stage1_train_B8_96x96_full_head": (8, 4, 96, 96, dict(free_residual=True, clamp_flow_t=20.0), 50),
        "stv2_stage1_train_B8_48x48_affine_full_head": (8, 4, 48, 48, dict(free_residual_with_affine=True, clamp_flow_t=20.0), 50),
    }
    for name, (B_, K_, H_, W_, kw, steps) in cases.items():
        try:
            ins = synthetic_inputs(B_, K_, H_, W_, seed=0, device=dev)
            imgs = torch.zeros(B_, 2, 3, 8, 8)
            res = {}
            for impl in ("ours", "torch_eager_port"):
                torch.manual_seed(1)
                if impl == "ours":
                    head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, mask_layer=K_,
                                                               mask_size=(H_, W_), **kw).to(dev)
                else:
                    head = PortedHead(mask_layer=K_, mask_size=(H_, W_), **kw).to(dev)
                m = ins[0].clone().requires_grad_(True)
                r1 = ins[3].clone().requires_grad_(True)
                r2 = ins[4].clone().requires_grad_(True)
                # parameter gradients only for real heads: cuDNN's wgrad for the degenerate 2->2 1x1 proxy convs
                # over 6.5M pixels takes ~16 ms in BOTH arms and would hide everything else
                params = list(head.parameters()) if kw.get("num_flow_feat_channels", 64) > 2 else []

                def fn():
                    _, l = head(imgs, m, ins[1], ins[2], r1, r2)
                    torch.autograd.grad(l["seg"], [m, r1, r2, *params])

                ms = _time_cuda(fn, steps)
                res[impl] = {"ms_per_step": ms, "samples_per_s": B_ / ms * 1e3}
                if impl == "ours" and H_ * W_ <= 128 * 128:
                    # launch-bound regime: the same head captured once into CUDA graphs (fwd graph + bwd graph)
                    try:
                        from rcf_unsupvideoseg_b200.graphed import make_graphed_head
                        gh = make_graphed_head(head, (imgs, m, ins[1], ins[2], r1, r2))

                        def gfn():
                            l = gh(m, ins[1], ins[2], r1, r2)
                            torch.autograd.grad(l["seg"], [m, r1, r2, *params])

                        ms_g = _time_cuda(gfn, steps)
                        res["ours_cuda_graph"] = {"ms_per_step": ms_g, "samples_per_s": B_ / ms_g * 1e3}
                    except Exception as ex:  # noqa: BLE001
                        res["ours_cuda_graph"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
                del head
                torch.cuda.empty_cache()
            res["speedup_vs_torch_eager"] = res["torch_eager_port"]["ms_per_step"] / res["ours"]["ms_per_step"]
            out[name] = res
        except Exception as ex:  # noqa: BLE001
            out[name] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
            torch.cuda.empty_cache()
    out.update(caller_side_step(pkg, dev))
    return out


def caller_side_step(pkg, dev):
    """The caller-side slice of RCFModel.forward_train around the head at the DAVIS stage-1 training shape
    (configs/rcf/rcf_stage1.yaml: B=8, mask_size 96x96, residual predicted at 48x48, w_seg 1, w_entropy 0.05):
    logits -> softmax (+ log_softmax of it, entropy loss) -> head (residual resize 48->96 inside) -> total loss ->
    gradients w.r.t. logits, residuals and head parameters.  ours = mask_losses + drop-in head; baseline = the ATen op
    sequence of models/rcf_model.py:433-434, :376-378 + the op-for-op port of the head, eager on the same GPU."""
    import torch
    import torch.nn.functional as F

    from oracle.torch_port import PortedHead
    from rcf_unsupvideoseg_b200.mask_ops import mask_losses
    name = "davis_stage1_caller_side_step_B8_96x96_resid48"
    try:
        B_, K_, H_, W_ = 8, 4, 96, 96
        g = torch.Generator(device=dev).manual_seed(0)
        logits = (torch.randn(B_, 2, K_, H_, W_, device=dev, generator=g) * 2).requires_grad_(True)
        fw = torch.randn(B_, 1, 2, H_, W_, device=dev, generator=g) * 8
        bw = torch.randn(B_, 1, 2, H_, W_, device=dev, generator=g) * 8
        r1 = (torch.randn(B_, 2 * K_, H_ // 2, W_ // 2, device=dev, generator=g) * 5).requires_grad_(True)
        r2 = (torch.randn(B_, 2 * K_, H_ // 2, W_ // 2, device=dev, generator=g) * 5).requires_grad_(True)
        imgs = torch.zeros(B_, 2, 3, 8, 8)
        kw = dict(mask_layer=K_, mask_size=(H_, W_), free_residual=True, clamp_flow_t=20.0, allow_residual_resize=True)
        res = {}
        for impl in ("ours", "torch_eager_port"):
            torch.manual_seed(1)
            if impl == "ours":
                head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, **kw).to(dev)
                head.return_flows = False
            else:
                head = PortedHead(**kw).to(dev)
            params = list(head.parameters())

            def fn():
                if impl == "ours":
                    masks, ml = mask_losses(logits)
                    ent = ml["entropy"]
                else:
                    masks = F.softmax(logits, dim=2)
                    ent = -(masks * F.log_softmax(masks, dim=2)).sum(dim=2).mean()
                _, l = head(imgs, masks, fw, bw, r1, r2)
                torch.autograd.grad(l["seg"] + 0.05 * ent, [logits, r1, r2, *params])

            ms = _time_cuda(fn, 50)
            res[impl] = {"ms_per_step": ms, "samples_per_s": B_ / ms * 1e3}
            if impl == "ours":
                try:      # the same step replayed from one CUDA graph
                    import gc
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        fn()
                    torch.cuda.current_stream().wait_stream(side)
                    gc.collect()
                    torch.cuda.synchronize()
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg):
                        fn()
                    ms_g = _time_cuda(cg.replay, 50)
                    res["ours_cuda_graph"] = {"ms_per_step": ms_g, "samples_per_s": B_ / ms_g * 1e3}
                except Exception as ex:  # noqa: BLE001
                    res["ours_cuda_graph"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
                    torch.cuda.synchronize()
            del head, params
            torch.cuda.empty_cache()
        res["speedup_vs_torch_eager"] = res["torch_eager_port"]["ms_per_step"] / res["ours"]["ms_per_step"]
        return {name: res}
    except Exception as ex:  # noqa: BLE001
        torch.cuda.empty_cache()
        return {name: {"error": f"{type(ex).__name__}: {ex}"[:200]}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the add-on measurements (full head, eager PyTorch on the same GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
