"""Generate tests/golden/warp_ops.npz and loss_utils.npz from the UNMODIFIED reference files
utils/warp_utils.py and utils/loss_utils.py (build container only; they import nothing but torch)."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    wu = load("/root/reference/utils/warp_utils.py", "ref_warp_utils")
    lu = load("/root/reference/utils/loss_utils.py", "ref_loss_utils")
    g = torch.Generator().manual_seed(7)
    out = {}
    B, C, H, W = 2, 3, 13, 17
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    flow = torch.randn(B, 2, H, W, generator=g, dtype=torch.float64) * 3.0
    flow[0, :, 0, :4] = 0.0                 # integer positions incl. the frame corner
    flow[1, 0, :, -1] = 2.5                 # push the last column out of the frame
    gout = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    out["x"], out["flow"], out["gout"] = x.numpy(), flow.numpy(), gout.numpy()
    for pad in ("border", "zeros"):
        xr = x.clone().requires_grad_(True); fr = flow.clone().requires_grad_(True)
        y = wu.flow_warp(xr, fr, pad=pad)
        y.backward(gout)
        out[f"warp.{pad}"] = y.detach().numpy()
        out[f"warp.{pad}.gx"] = xr.grad.numpy()
        out[f"warp.{pad}.gflow"] = fr.grad.numpy()
    flow21 = torch.randn(B, 2, H, W, generator=g, dtype=torch.float64) * 2.0
    out["flow21"] = flow21.numpy()
    grid = wu.mesh_grid(B, H, W).type_as(flow21) + flow21
    out["corr_map"] = wu.get_corresponding_map(grid).numpy()
    out["occ_backward"] = wu.get_occu_mask_backward(flow21, th=0.2).numpy()
    out["occ_bidirection"] = wu.get_occu_mask_bidirection(flow, flow21).numpy()
    out["mesh_grid"] = wu.mesh_grid(2, 3, 4).numpy()
    out["norm_grid"] = wu.norm_grid(wu.mesh_grid(2, 3, 4).double()).numpy()
    np.savez_compressed(os.path.join(HERE, "aux", "warp_ops.npz"), **out)

    lo = {}
    p = torch.softmax(torch.randn(2, 4, 5, 6, generator=g, dtype=torch.float64), 1)
    lo["p"] = p.numpy()
    lo["sharpen_T0.25"] = lu.sharpen(p, 0.25, dim=1).numpy()
    loss = torch.rand(3, 4, 5, generator=g, dtype=torch.float64)
    w = torch.rand(3, 1, 5, generator=g, dtype=torch.float64)
    lo["loss"], lo["weight"] = loss.numpy(), w.numpy()
    lo["wrl.mean"] = lu.weight_reduce_loss(loss, w, "mean").numpy()
    lo["wrl.sum"] = lu.weight_reduce_loss(loss, w, "sum").numpy()
    lo["wrl.none"] = lu.weight_reduce_loss(loss, w, "none").numpy()
    lo["wrl.avg"] = lu.weight_reduce_loss(loss, w, "mean", avg_factor=7.0).numpy()
    np.savez_compressed(os.path.join(HERE, "aux", "loss_utils.npz"), **lo)
    print("written", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
