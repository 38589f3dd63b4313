"""Worker of tests/test_ddp_nccl.py (one process per GPU, launched by torchrun): the drop-in head wrapped in
DistributedDataParallel on a batch shard, checked on rank 0 against the single-GPU run over the global batch."""
import os
import sys

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcf_unsupvideoseg_b200 as pkg  # noqa: E402


def inputs(B, K, H, W, seed, dev):
    g = torch.Generator().manual_seed(seed)
    masks = torch.softmax(torch.randn(B, 2, K, H, W, generator=g) * 2.0, dim=2)
    fw = torch.randn(B, 1, 2, H, W, generator=g) * 8.0
    bw = torch.randn(B, 1, 2, H, W, generator=g) * 8.0
    rfw = torch.randn(B, 2 * K, H, W, generator=g) * 5.0
    rbw = torch.randn(B, 2 * K, H, W, generator=g) * 5.0
    return [t.to(dev) for t in (masks, fw, bw, rfw, rbw)]


def run(head, masks, fw, bw, rfw, rbw):
    masks = masks.clone().requires_grad_(True); rfw = rfw.clone().requires_grad_(True); rbw = rbw.clone().requires_grad_(True)
    for p in head.parameters():
        p.grad = None
    imgs = torch.zeros(masks.shape[0], 2, 3, 8, 8, device=masks.device)
    _, loss = head(imgs, masks, fw, bw, rfw, rbw)
    loss["seg"].backward()
    return loss["seg"].detach(), masks.grad, rfw.grad, rbw.grad


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda")
    dist.init_process_group("nccl")
    torch.backends.cudnn.allow_tf32 = False
    affine = "--affine" in sys.argv
    B, K, H, W = 2 * world, 4, 64, 96
    kw = dict(free_residual_with_affine=True) if affine else dict(free_residual=True)
    torch.manual_seed(1)
    head = pkg.FlowAggregationHeadWithResidual(args=None, create_flownet=True, mask_layer=K, mask_size=(H, W),
                                               clamp_flow_t=20.0, **kw).to(dev)
    head.return_flows = False
    full = inputs(B, K, H, W, 3, dev)
    lo, hi = rank * (B // world), (rank + 1) * (B // world)
    ddp = DDP(head, device_ids=[dev.index])
    masks = full[0][lo:hi].clone().requires_grad_(True)
    rfw = full[3][lo:hi].clone().requires_grad_(True)
    rbw = full[4][lo:hi].clone().requires_grad_(True)
    imgs = torch.zeros(hi - lo, 2, 3, 8, 8, device=dev)
    _, loss = ddp(imgs, masks, full[1][lo:hi], full[2][lo:hi], rfw, rbw)
    loss["seg"].backward()                      # DDP all-reduces (averages) the 42 434 parameter gradients over NCCL
    lsum = loss["seg"].detach().clone()
    dist.all_reduce(lsum)
    ddp_grads = {k: p.grad.clone() for k, p in head.named_parameters()}
    ok = True
    # single-GPU reference over the global batch (every rank computes it; rank 0 reports)
    l_ref, dm_ref, dfw_ref, dbw_ref = run(head, *full)
    ref_grads = {k: p.grad.clone() for k, p in head.named_parameters()}
    errs = {"loss": abs(float(lsum) / world - float(l_ref)) / float(l_ref),
            "d_masks": rel(masks.grad / world, dm_ref[lo:hi]), "d_resid_fw": rel(rfw.grad / world, dfw_ref[lo:hi]),
            "d_resid_bw": rel(rbw.grad / world, dbw_ref[lo:hi])}
    errs.update({"dparam." + k: rel(ddp_grads[k], ref_grads[k]) for k in ddp_grads})
    tol = {k: (1e-5 if k == "loss" else 1e-4 if k.startswith("d_") else 3e-4) for k in errs}
    bad = {k: v for k, v in errs.items() if not v <= tol[k]}
    flag = torch.tensor([1.0 if bad else 0.0], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("DDP_NCCL_ERRS", {k: f"{v:.2e}" for k, v in errs.items()}, flush=True)
        print("DDP_NCCL_OK" if float(flag) == 0 else f"DDP_NCCL_FAIL {bad}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if float(flag) == 0 else 1)


if __name__ == "__main__":
    main()
