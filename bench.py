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
  e2e    : the drop-in nn.Module call (FlowAggregationHeadWithResidual with the reference's DEFAULT constructor:
           64 feature channels, 3x3 convs -- conv stem, tcgen05 convs, pooling, MLP, loss, all in librcf_loss.so) fed
           from PINNED HOST buffers every step: H2D of masks/flows/residuals, fwd+bwd, D2H of the loss scalars and
           of the mask/residual gradients.
  roofline: dominant kernel k_bwd (streaming backward): algorithmic bytes per launch / mean CUDA-event
           duration of that kernel inside the timed region (library timing hook), vs MEASURED_PEAKS.json.
  cpu_baseline: oracle/torch_port.py (op-for-op PyTorch CPU port of the reference head, pinned to the
           reference's outputs by tests; same default constructor) on the host cores, bounded sample (B=2), rank 0, N=1.
  extras : the rest of BASELINE.json's configs in the same record: loss-core sweep (C2 affine, C4 K=2/4/8, C5, FBMS),
           C3 as written (global B=64 split over the ranks, free + affine), the drop-in module device-resident
           (value_module), the tcgen05 conv kernels next to cuDNN, training shapes, PyTorch-eager port on the same GPU.
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
HEAD_KW = dict(mask_layer=K, mask_size=(H, W), free_residual=True, clamp_flow_t=20.0)     # default head: Cf=64, 3x3 convs
NBLOCKS = 5          # timed blocks of `steps` steps each; the median block is reported


def synthetic_inputs(B, K_, H_, W_, seed, device=None):
    """SURVEY.md 8(d) synthetic inputs from torch's CPU generator (bit-reproducible on any machine)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    masks = torch.softmax(torch.randn(B, 2, K_, H_, W_, generator=g) * 2.0, dim=2)
    fw = torch.randn(B, 1, 2, H_, W_, generator=g) * 8.0
    bw = torch.randn(B, 1, 2, H_, W_, generator=g) * 8.0
    rfw = torch.randn(B, 2 * K_, H_, W_, generator=g) * 5.0
    rbw = torch.randn(B, 2 * K_, H_, W_, generator=g) * 5.0
    out = (masks, fw, bw, rfw, rbw)
    return tuple(t.to(device) for t in out) if device is not None else out


def workload_config(n_gpus):
    return {
        "workload": f"C2: RCF loss fwd+bwd, B={B_PER_GPU}/GPU, K={K}, {H}x{W}, free_residual+L1+clamp20 (stage-1 DAVIS flags)",
        "value_path": "loss core, theta supplied (SURVEY 8d Scope L), C ABI rcf_forward+rcf_backward, device-resident",
        "e2e_path": "drop-in head, default constructor (Cf=64, 3x3 convs: stem + tcgen05 convs + pooling + MLP + loss), "
                    "pinned host buffers in, loss+grads out",
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

    from oracle.torch_port import PortedHead
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
            "sample": f"B={Bs} of the C2 workload ({K=}, {H}x{W}, default head Cf=64 3x3), {len(times)} timed steps, "
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
    numa_note = None
    if world > 1:
        # e2e is host-memory / PCIe bound: bind this rank to the CPUs of its GPU's NUMA node BEFORE the pinned staging
        # buffers are allocated (first touch puts their pages next to the GPU's root port)
        try:
            import pynvml
            pynvml.nvmlInit()
            h_ = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            pynvml.nvmlDeviceSetCpuAffinity(h_)
            numa_note = f"rank bound to the {len(os.sched_getaffinity(0))} CPUs NVML reports as local to GPU {local_rank}"
        except Exception as ex:  # noqa: BLE001
            numa_note = f"NUMA binding unavailable ({type(ex).__name__})"
        # keep stdout clean for the one JSON line: NCCL's banner / debug output goes to a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/rcf_bench_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)
    lib = pkg.load_library(build_if_missing=False)

    B = B_PER_GPU
    P = H * W
    inv_n = 1.0 / (B * world * 2 * P)
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
        block_ms = []
        for _blk in range(NBLOCKS):          # NBLOCKS regions of exactly `steps` steps, each bracketed by barrier + sync
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
            tb = torch.tensor([gs_.elapsed_time(ge_) / args.steps], device=dev)
            if world > 1:
                dist.all_reduce(tb, op=dist.ReduceOp.MAX)      # max over ranks, per block
            block_ms.append(float(tb))
        graph_ms = sorted(block_ms)[len(block_ms) // 2]
        lfin = loss_g.detach().clone()
        if world > 1:
            dist.all_reduce(lfin)
        loss_val = [float(x) for x in lfin.cpu()]
    except Exception as ex:  # noqa: BLE001
        print(f"[bench] CUDA graph capture unavailable ({type(ex).__name__}: {str(ex)[:300]}); using eager timing",
              file=sys.stderr)
        torch.cuda.synchronize()
    t = torch.tensor([eager_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    eager_ms = float(t[0])
    if graph_ms is None:
        block_ms, graph_ms = [], -1.0
    ms_step = eager_ms if graph_ms <= 0 else min(graph_ms, eager_ms)
    value = B * world / (ms_step * 1e-3)
    kb_ms = sorted(a.elapsed_time(b) for a, b in ev_pairs)
    kb_mean = sum(kb_ms) / len(kb_ms)

    # e2e through the drop-in module, pinned host buffers in / loss + grads out
    torch.manual_seed(1)
    head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, **HEAD_KW).to(dev)
    head.return_flows = False
    head.loss_inv_n = inv_n
    wc_note = None
    if os.environ.get("RCF_BENCH_WC", "0") == "1":
        # experiment: write-combined pinned staging buffers for the host->device sources (the CPU only ever writes them)
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        pin = []
        for t in (masks_h, fw_h, bw_h, rfw_h, rbw_h):
            ptr = ctypes.c_void_p()
            rc_ = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(t.numel() * 4), ctypes.c_uint(0x04))   # cudaHostAllocWriteCombined
            assert rc_ == 0, f"cudaHostAlloc -> {rc_}"
            buf = (ctypes.c_float * t.numel()).from_address(ptr.value)
            w_ = torch.frombuffer(buf, dtype=torch.float32).view(t.shape)
            w_.copy_(t)
            pin.append(w_)
        wc_note = "host->device staging buffers are write-combined pinned memory (cudaHostAllocWriteCombined)"
    else:
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

    # The host's ceiling for this traffic pattern: the same host->device and device->host copies on the same two streams
    # with NO compute between them, all ranks at once (what the PCIe / host-memory fabric gives this many concurrent GPUs).
    d_out_src = [torch.empty_like(t, device=dev) for t in out_pin]

    def copy_only_step(i):
        with torch.cuda.stream(s_in):
            for d, src in zip(d_in[i % NBUF], pin):
                d.copy_(src, non_blocking=True)
        with torch.cuda.stream(s_out):
            for dst, src in zip(out_pin, d_out_src):
                dst.copy_(src, non_blocking=True)

    for i in range(2):
        copy_only_step(i)
    s_cmp.wait_stream(s_in); s_cmp.wait_stream(s_out)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    s_in.wait_event(c0); s_out.wait_event(c0)
    for i in range(e2e_steps):
        copy_only_step(i)
    s_cmp.wait_stream(s_in); s_cmp.wait_stream(s_out)
    c1.record()
    torch.cuda.synchronize()
    barrier()
    tc = torch.tensor([c0.elapsed_time(c1)], device=dev)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    copy_only_ms = float(tc) / e2e_steps
    del d_out_src

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
                    "note": "PCIe-bound; copies of neighbouring steps overlap compute on 3 streams", "host_numa": numa_note,
                    "copy_only_ms_per_step": copy_only_ms,
                    "copy_only_gbs_each_way": [h2d / copy_only_ms / 1e6, d2h / copy_only_ms / 1e6],
                    "copy_only_note": "the same H2D + D2H copies with no compute, all ranks at once: the host's ceiling for e2e",
                    "staging": wc_note or "torch pin_memory() (cudaHostAlloc default flags)",
                    "conv_precision": "follows torch.backends.cudnn.allow_tf32 (torch default True -> one product of fp16 "
                                      "operands, TF32-class; False -> 3 bf16 products per fp32 product, fp32-grade)"},
            # value region: k_loss, k_finalize, k_loss_sum | k_segment_bwd, k_bwd per step (single-pass forward), NBLOCKS blocks;
            # e2e region: the library's 19 kernels of the default head per step (profiles/r02_timelines_final.md lists 21: one more,
            # k_bias_grad_final, was folded into k_bias_grad_fd afterwards) + autograd's one-element fill of the upstream gradient)
            "gpu_launches": 5 * args.steps * NBLOCKS + 19 * e2e_steps,
            "roofline": {"bound": "hbm", "kernel": "k_bwd<K=4,D=0,PX=4> (streaming backward)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": read_traffic("k_bwd<4, 0, 4>"),
                         "peak_source": peak_src, "kernel_ms": kb_mean, "kernel_ms_min": kb_ms[0],
                         "algorithmic_bytes_per_launch": kb_bytes},
            "step_roofline": {"algorithmic_gbs": step_gbs, "frac": step_gbs / peak,
                              "frame_directions_per_s": 2 * value},
            "timing": {"eager_ms_per_step": eager_ms, "cuda_graph_ms_per_step": (graph_ms if graph_ms > 0 else None),
                       "cuda_graph_block_ms": block_ms, "blocks": f"{NBLOCKS} x {args.steps} steps, median block reported",
                       "cpu_enqueue_ms_per_step": cpu_enqueue_ms,
                       "value_from": "cuda_graph" if (graph_ms > 0 and graph_ms <= eager_ms) else "eager"},
            "loss": loss_val,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _, _ = cpu_port_time(8, 1, budget_s=25.0)
        line["extras"] = {}
    c3 = None if args.no_extras else c3_strong_scaling(pkg, dev, world, rank)
    if rank == 0:
        if c3 is not None:
            line["extras"]["c3_strong_scaling_global_B64"] = c3
        if world == 1 and not args.no_extras:
            line["extras"].update(extras(pkg, dev))
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



def _graph_time(step, steps, blocks=3, sync=None):
    """Captures `step` (a fwd+bwd closure) into one CUDA graph and times `blocks` x `steps` replays; median block, ms/step."""
    import gc

    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    gc.collect()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = step()
    for _ in range(3):
        gr.replay()
    res = []
    for _ in range(blocks):
        if sync is not None:
            sync()
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(steps):
            gr.replay()
        e_.record()
        torch.cuda.synchronize()
        res.append(s_.elapsed_time(e_) / steps)
    del out, gr
    torch.cuda.empty_cache()
    return sorted(res)[len(res) // 2]


def _core_step(pkg, dev, B, K_, H_, W_, D, robust=False, inv_n=0.0, seed=0):
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    masks = torch.softmax(torch.randn(B, 2, K_, H_, W_, device=dev, generator=g) * 2, dim=2).requires_grad_(True)
    flows = [torch.randn(B, 2, H_, W_, device=dev, generator=g) * 8 for _ in range(2)]
    resids = [(torch.randn(B, 2 * K_, H_, W_, device=dev, generator=g) * 5).requires_grad_(True) for _ in range(2)]
    thetas = [torch.randn(B, 2, K_, device=dev, generator=g).requires_grad_(True) for _ in range(2)]
    spec = pkg.LossSpec(K=K_, H=H_, W=W_, D=D, Cf=0, clamp_t=20.0, robust=robust, inv_n=inv_n)
    gl = torch.ones(2, device=dev)

    def step():
        loss, _ = pkg.rcf_motion_loss(spec, masks, flows, resids, thetas=thetas)
        return loss, torch.autograd.grad(loss, [masks, *resids, *thetas], grad_outputs=gl)
    return step


def _module_step(pkg, dev, B, K_, H_, W_, kw, inv_n=0.0, nprod=None, seed=0):
    import torch
    torch.manual_seed(1)
    head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, mask_layer=K_, mask_size=(H_, W_), **kw).to(dev)
    head.return_flows = False
    head.loss_inv_n = inv_n
    head.conv_precision = nprod
    ins = synthetic_inputs(B, K_, H_, W_, seed=seed, device=dev)
    m = ins[0].requires_grad_(True)
    r1 = ins[3].requires_grad_(True)
    r2 = ins[4].requires_grad_(True)
    imgs = torch.zeros(B, 2, 3, 8, 8)
    params = list(head.parameters())

    def step():
        _, l = head(imgs, m, ins[1], ins[2], r1, r2)
        return torch.autograd.grad(l["seg"], [m, r1, r2, *params])
    return step


def loss_core_sweep(pkg, dev):
    """BASELINE.json configs C2 (affine), C4 (K = 2/4/8, B = 32), C5 (three resolutions) and the FBMS K = 3 affine flags
    through the loss core (theta supplied), CUDA-graph replay; `frac` = SURVEY 8(d) algorithmic bytes (36K+16 B/px/dir)
    / time / measured HBM copy rate.  Affine rows need pass 1 (4K+8 B/px/dir more traffic: ceiling 160/184 = 87 % at K=4)."""
    peak, _ = read_peaks()
    rows = [("C2_free", 16, 4, 480, 854, 0), ("C2_affine", 16, 4, 480, 854, 2), ("C4_K2_free", 32, 2, 480, 854, 0),
            ("C4_K4_free", 32, 4, 480, 854, 0), ("C4_K8_free", 32, 8, 480, 854, 0), ("C4_K8_affine", 32, 8, 480, 854, 2),
            ("FBMS_K3_affine", 16, 3, 480, 854, 2), ("C5_240x427_free", 16, 4, 240, 427, 0),
            ("C5_480x854_free", 16, 4, 480, 854, 0), ("C5_1080x1920_free", 16, 4, 1080, 1920, 0),
            ("DAVIS_train_96x96_B8_free", 8, 4, 96, 96, 0), ("STv2_train_48x48_B8_affine", 8, 4, 48, 48, 2)]
    out = {}
    for name, B, K_, H_, W_, D in rows:
        ms = _graph_time(_core_step(pkg, dev, B, K_, H_, W_, D), 20)
        alg = B * 2 * H_ * W_ * (36 * K_ + 16)
        out[name] = {"B": B, "K": K_, "HxW": f"{H_}x{W_}", "mode": {0: "free", 2: "affine"}[D], "ms_per_step": ms,
                     "samples_per_s": B / ms * 1e3, "algorithmic_gbs": alg / ms / 1e6, "frac": alg / ms / 1e6 / peak}
    return out


def module_throughput(pkg, dev):
    """The drop-in nn.Module (default head: Cf = 64, 3x3 convs) device-resident at C2 (B = 16) and C1 (B = 2), fwd + bwd
    incl. all 8 parameter gradients, CUDA-graph replay, for each conv precision."""
    out = {}
    for name, B in (("C2_B16", 16), ("C1_B2", 2)):
        for nprod in (2, 3, 1):
            ms = _graph_time(_module_step(pkg, dev, B, K, H, W, dict(free_residual=True, clamp_flow_t=20.0), nprod=nprod), 5)
            out[f"{name}_nprod{nprod}"] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3}
    out["note"] = ("nprod = conv precision of the head: 3 fp32-grade (three bf16 products per fp32 product; "
                   "torch.backends.cudnn.allow_tf32 = False), 2 TF32-class (ONE product of fp16 operands, the feature-map gradient "
                   "carried as scaled fp16; torch's default and inside autocast), 1 plain bf16 (conv_precision = 1 only)")
    return out


def conv_kernels(pkg, dev):
    """The three tcgen05 conv kernels against the cuDNN kernels they replace, same tensors (4 x 64 x 480 x 854 = C1)."""
    import torch
    import torch.nn.functional as F

    from rcf_unsupvideoseg_b200 import conv64 as c64
    N_ = 4
    x = torch.randn(N_, 64, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(N_, 64, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, 64, 3, 3, device=dev) / 24
    wp, wpt = c64.pack_weights(w, False), c64.pack_weights(w, True)
    hi, lo = c64.split_bf16(x)
    g_hi, g_lo = c64.split_bf16(gy)
    flops = 2.0 * N_ * H * W * 64 * 64 * 9
    burst, sustained = 1630.8, 1391.4
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
            burst, sustained = float(pk["bf16_tflops"]), float(pk["bf16_tflops_sustained"])
    except Exception:  # noqa: BLE001
        pass
    out = {"gflop_per_kernel": flops / 1e9,
           "tensor_roofline": {"bound": "tensor", "peak_burst_tflops": burst, "peak_sustained_tflops": sustained, "unit": "TFLOP/s",
                               "definition": "bf16 FLOPs the kernel issues to the tensor cores (nprod products per fp32 product; the weight "
                                             "gradient also computes the centre tap row twice: x 12/9) / CUDA-event time / the measured "
                                             "cuBLAS bf16 rate (MEASURED_PEAKS.json); halo rows (4-5 %) not counted"}}
    out["kernel_modes"] = ("kernel-level nprod: 1 = one product per tap (bf16 operands; the head's TF32-class default runs this same "
                           "instruction stream with fp16 operand formats), 2 = activations x [W_hi | W_lo] stacked in N = 128 "
                           "(no longer used by the head), 3 = three products (fp32-grade)")
    for nprod in (1, 2, 3):
        t_f = _time_cuda(lambda: c64.conv64_pair(hi, lo, wp, nprod), 10) * 1e3
        t_d = _time_cuda(lambda: c64.conv64_pair(g_hi, g_lo, wpt, nprod), 10) * 1e3
        t_w = _time_cuda(lambda: c64.conv64_wgrad_pair(hi, lo, g_hi, g_lo, nprod), 10) * 1e3
        ex_f, ex_w = nprod * flops / t_f / 1e6, nprod * (12.0 / 9.0) * flops / t_w / 1e6
        out[f"tcgen05_nprod{nprod}_us"] = {"fprop": t_f, "dgrad": t_d, "wgrad": t_w, "fprop_useful_fp32_tflops": flops / t_f / 1e6,
                                           "fprop_executed_bf16_tflops": ex_f, "fprop_frac_of_burst_peak": ex_f / burst,
                                           "fprop_frac_of_sustained_peak": ex_f / sustained,
                                           "wgrad_executed_bf16_tflops": ex_w, "wgrad_frac_of_burst_peak": ex_w / burst,
                                           "wgrad_frac_of_sustained_peak": ex_w / sustained}
    old = torch.backends.cudnn.allow_tf32
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        cb = lambda mask: torch.ops.aten.convolution_backward(gy, x, w, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, mask)  # noqa: E731
        out[f"cudnn_{'tf32' if tf32 else 'fp32'}_us"] = {
            "fprop": _time_cuda(lambda: F.conv2d(x, w, None, 1, 1), 5) * 1e3,
            "dgrad": _time_cuda(lambda: cb((True, False, False)), 5) * 1e3,
            "wgrad": _time_cuda(lambda: cb((False, True, False)), 5) * 1e3}
    torch.backends.cudnn.allow_tf32 = old
    return out


def c3_strong_scaling(pkg, dev, world, rank):
    """BASELINE.json config C3 as written: global batch 64 at K = 4, 480x854, split over the ranks (64 / N per GPU:
    strong scaling), free_residual and free_residual_with_affine, loss core and drop-in module.  The mean's normaliser is
    the global element count, so per-rank losses sum to the global loss (one 2-float NCCL all-reduce per step, outside
    the timed graph).  Times are the max over ranks."""
    import torch
    import torch.distributed as dist
    GB = 64
    if GB % world:
        return None
    Bl = GB // world
    inv_n = 1.0 / (GB * 2 * H * W)
    peak, _ = read_peaks()

    def sync():
        if world > 1:
            dist.barrier()

    out = {"global_batch": GB, "per_gpu_batch": Bl, "n_gpus": world, "scaling": "strong"}
    for name, mk in (("loss_core_free", lambda: _core_step(pkg, dev, Bl, K, H, W, 0, inv_n=inv_n, seed=rank)),
                     ("loss_core_affine", lambda: _core_step(pkg, dev, Bl, K, H, W, 2, inv_n=inv_n, seed=rank)),
                     ("module_free_default_head", lambda: _module_step(pkg, dev, Bl, K, H, W, dict(free_residual=True, clamp_flow_t=20.0),
                                                                       inv_n=inv_n, seed=rank)),
                     ("module_affine_default_head", lambda: _module_step(pkg, dev, Bl, K, H, W,
                                                                         dict(free_residual_with_affine=True, clamp_flow_t=20.0),
                                                                         inv_n=inv_n, seed=rank))):
        try:
            ms = _graph_time(mk(), 10 if "core" in name else 3, sync=sync)
        except Exception as ex:  # noqa: BLE001
            ms = float("nan")
            print(f"[bench] c3 {name}: {type(ex).__name__}: {str(ex)[:200]}", file=sys.stderr)
            torch.cuda.empty_cache()
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        rec = {"ms_per_step": ms, "samples_per_s": GB / ms * 1e3}
        if "core" in name:
            rec["frac_of_hbm_peak_per_gpu"] = Bl * 2 * H * W * (36 * K + 16) / ms / 1e6 / peak
        out[name] = rec
    return out


def extras(pkg, dev):
    """Add-on measurements reported beside the headline (SURVEY.md 8(d) 'report separately'):
    the same workloads through (a) the drop-in head and (b) the op-for-op PyTorch port of the reference
    running EAGER ON THE SAME GPU (the GPU baseline the reference user has today)."""
    import torch

    from oracle.torch_port import PortedHead       # baseline arm only: the reference's op sequence, eager on this GPU
    out = {}
    try:
        out["loss_core_sweep"] = loss_core_sweep(pkg, dev)
    except Exception as ex:  # noqa: BLE001
        out["loss_core_sweep"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
        torch.cuda.empty_cache()
    try:
        out["value_module"] = module_throughput(pkg, dev)
    except Exception as ex:  # noqa: BLE001
        out["value_module"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
        torch.cuda.empty_cache()
    try:
        out["conv64_kernels_vs_cudnn_4x64x480x854"] = conv_kernels(pkg, dev)
    except Exception as ex:  # noqa: BLE001
        out["conv64_kernels_vs_cudnn_4x64x480x854"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
        torch.cuda.empty_cache()
    cases = {
        "c2_proxy_head_B16_480x854": (16, 4, 480, 854, dict(free_residual=True, clamp_flow_t=20.0, num_flow_feat_channels=2,
                                                             flow_feat_before_agg_kernel_size=1), 10),
        "c1_full_head_B2_480x854": (2, 4, 480, 854, dict(free_residual=True, clamp_flow_t=20.0), 10),
        "davis_stage1_train_B8_96x96_full_head": (8, 4, 96, 96, dict(free_residual=True, clamp_flow_t=20.0), 50),
        "stv2_stage1_train_B8_48x48_affine_full_head": (8, 4, 48, 48, dict(free_residual_with_affine=True, clamp_flow_t=20.0), 50),
        # the AMP configs (configs/rcf_stv2/rcf_stage1.yaml:60, configs/rcf_fbms59/rcf_stage1.yaml:61 `precision: 16`):
        # BOTH arms inside torch.autocast(fp16) -- the port runs its convs / einsums in fp16 and the solve in fp32 like
        # the reference (:215-217); the drop-in runs the tcgen05 convs with fp16 operands (level 2) and the loss core in fp32
        "amp_stv2_stage1_train_B8_48x48_affine_full_head": (8, 4, 48, 48, dict(free_residual_with_affine=True, clamp_flow_t=20.0), 50),
        "amp_fbms_K3_B2_480x854_affine_full_head": (2, 3, 480, 854, dict(free_residual_with_affine=True, clamp_flow_t=20.0), 10),
    }
    import contextlib
    for name, (B_, K_, H_, W_, kw, steps) in cases.items():
        amp = (lambda: torch.autocast("cuda", dtype=torch.float16)) if name.startswith("amp_") else contextlib.nullcontext
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
                    with amp():
                        _, l = head(imgs, m, ins[1], ins[2], r1, r2)
                    torch.autograd.grad(l["seg"].float(), [m, r1, r2, *params])

                ms = _time_cuda(fn, steps)
                res[impl] = {"ms_per_step": ms, "samples_per_s": B_ / ms * 1e3}
                if impl == "ours" and H_ * W_ <= 128 * 128 and not name.startswith("amp_"):
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
    try:
        out["warp_utils_kernels"] = warp_kernels(pkg, dev)
    except Exception as ex:  # noqa: BLE001
        out["warp_utils_kernels"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
        torch.cuda.empty_cache()
    return out


def warp_kernels(pkg, dev):
    """utils/warp_utils.py: flow_warp fwd+bwd and get_corresponding_map, ours vs the ATen op
    sequence the reference runs (grid_sample on a normalised grid; floor/clamp/cat + scatter_add_), same GPU, eager."""
    import torch
    import torch.nn.functional as F

    from rcf_unsupvideoseg_b200 import warp_utils as wu

    def ref_flow_warp(x, flow):                    # utils/warp_utils.py:84-94
        B_, _, H_, W_ = flow.shape
        ys, xs = torch.meshgrid(torch.arange(H_, device=dev, dtype=x.dtype), torch.arange(W_, device=dev, dtype=x.dtype), indexing="ij")
        gx = 2.0 * (xs[None] + flow[:, 0]) / (W_ - 1) - 1.0
        gy = 2.0 * (ys[None] + flow[:, 1]) / (H_ - 1) - 1.0
        return F.grid_sample(x, torch.stack([gx, gy], dim=-1), mode="bilinear", padding_mode="border", align_corners=True)

    def ref_corr_map(data):                        # utils/warp_utils.py:27-81
        B_, _, H_, W_ = data.shape
        x, y = data[:, 0].reshape(B_, -1), data[:, 1].reshape(B_, -1)
        xf, yf = torch.floor(x), torch.floor(y)
        acc = torch.zeros(B_, H_ * W_, device=dev, dtype=data.dtype)
        for cx, cy in ((xf + 1, yf + 1), (xf + 1, yf), (xf, yf + 1), (xf, yf)):
            cxc, cyc = cx.clamp(0, W_ - 1), cy.clamp(0, H_ - 1)
            v = (1 - (x - cxc).abs()) * (1 - (y - cyc).abs())
            v = torch.where((cx != cxc) | (cy != cyc), torch.zeros_like(v), v)
            acc.scatter_add_(1, (cxc + cyc * W_).long(), v)
        return acc.view(B_, 1, H_, W_)

    out = {}
    for tag, (B_, H_, W_) in {"B16_480x854": (16, 480, 854), "B8_96x96": (8, 96, 96)}.items():
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.randn(B_, 2, H_, W_, device=dev, generator=g).requires_grad_(True)
        flow = (torch.randn(B_, 2, H_, W_, device=dev, generator=g) * 4).requires_grad_(True)
        ys, xs = torch.meshgrid(torch.arange(H_, device=dev, dtype=torch.float32), torch.arange(W_, device=dev, dtype=torch.float32), indexing="ij")
        coords = torch.stack([xs, ys])[None] + flow.detach()
        go = torch.randn(B_, 2, H_, W_, device=dev, generator=g)
        row = {}
        for impl, fw_, cm_ in (("ours", lambda: wu.flow_warp(x, flow), lambda: wu.get_corresponding_map(coords)),
                               ("aten", lambda: ref_flow_warp(x, flow), lambda: ref_corr_map(coords))):
            def warp_step():
                torch.autograd.grad(fw_(), [x, flow], grad_outputs=go)
            row[impl] = {"flow_warp_fwd_bwd_ms": _time_cuda(warp_step, 20), "corresponding_map_ms": _time_cuda(cm_, 20)}
        row["speedup_flow_warp"] = row["aten"]["flow_warp_fwd_bwd_ms"] / row["ours"]["flow_warp_fwd_bwd_ms"]
        row["speedup_corresponding_map"] = row["aten"]["corresponding_map_ms"] / row["ours"]["corresponding_map_ms"]
        px = B_ * H_ * W_
        row["ours_flow_warp_algorithmic_gbs"] = px * (4 * 2 * 3 + 4 * 2 * 3) / row["ours"]["flow_warp_fwd_bwd_ms"] / 1e6   # fwd: x, flow, out; bwd: gout, gx, gflow (2 ch each)
        out[tag] = row
        del x, flow, coords, go
        torch.cuda.empty_cache()
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
