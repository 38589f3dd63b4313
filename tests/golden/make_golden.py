"""Generate tests/golden/*.npz by running the UNMODIFIED reference head (build container only).

    python tests/golden/make_golden.py            # rewrites every fixture

Each fixture stores the inputs (fp32), the head parameters, the reference's fp64 results (the
arbiter: same module, ``.double()``, forced ``.float()`` around linalg.solve neutralised) and the
reference's own fp32 results (what a user of the reference actually gets).  /root/reference does
not exist on the GPU box, so tests read only these files.

Also writes ``full_size_scalars.json``: loss scalars + gradient checksums of the reference (fp32)
at BASELINE.json's C1 shape (B=2, K=4, 480x854) for seeds reproducible from torch's CPU generator.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_loader  # noqa: E402

# name -> (shape dict, constructor kwargs)
CASES = {
    "free_l1": (dict(B=2, K=4, H=12, W=14), dict(free_residual=True, clamp_flow_t=20.0)),
    "free_robust": (dict(B=2, K=4, H=12, W=14), dict(num_flow_feat_channels=16, free_residual=True, clamp_flow_t=20.0, outlier_robust_loss=True)),
    "free_unbounded": (dict(B=2, K=4, H=12, W=14), dict(num_flow_feat_channels=16, free_residual=True, residual_adjustment_scale=-1.0)),
    "affine_l1": (dict(B=2, K=4, H=12, W=14), dict(free_residual_with_affine=True, clamp_flow_t=20.0)),
    "affine_robust": (dict(B=2, K=4, H=12, W=14), dict(num_flow_feat_channels=16, free_residual_with_affine=True, clamp_flow_t=20.0,
                                                         outlier_robust_loss=True, eps=0.02, q=0.5)),
    "quadratic_l1": (dict(B=2, K=3, H=12, W=14), dict(num_flow_feat_channels=16, free_residual_with_affine=True,
                                                        free_residual_with_affine_quadratic=True, clamp_flow_t=20.0)),
    "affine_k3_fbms": (dict(B=3, K=3, H=10, W=10), dict(num_flow_feat_channels=16, free_residual_with_affine=True, clamp_flow_t=20.0)),
    "free_resize": (dict(B=2, K=4, H=12, W=16, rh=6, rw=8), dict(num_flow_feat_channels=16, free_residual=True, clamp_flow_t=20.0,
                                                                  allow_residual_resize=True)),
    "affine_resize_odd": (dict(B=1, K=2, H=11, W=13, rh=5, rw=7), dict(num_flow_feat_channels=16, free_residual_with_affine=True,
                                                                        allow_residual_resize=True)),
    "free_norm_filter": (dict(B=2, K=5, H=9, W=11), dict(num_flow_feat_channels=16, free_residual=True, norm_flow=True, filter_flow_t=0.05,
                                                          clamp_flow_t=0.8, residual_adjustment_scale=4.0,
                                                          pred_div_coeff=5.0)),
    "free_k8_k1conv": (dict(B=1, K=8, H=8, W=12), dict(free_residual=True, clamp_flow_t=20.0,
                                                         flow_feat_before_agg_kernel_size=1,
                                                         num_flow_feat_channels=8)),
    "free_sharp_masks": (dict(B=2, K=4, H=12, W=14, sharp=8.0), dict(num_flow_feat_channels=16, free_residual=True, clamp_flow_t=20.0)),
}


def make_inputs(B, K, H, W, seed, rh=None, rw=None, sharp=1.0):
    """SURVEY.md 8(d) synthetic inputs from torch's CPU generator (bit-reproducible anywhere)."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, 2, K, H, W, generator=g) * 2.0 * sharp
    masks = torch.softmax(logits, dim=2)
    fw = torch.randn(B, 1, 2, H, W, generator=g) * 8.0
    bw = torch.randn(B, 1, 2, H, W, generator=g) * 8.0
    rh, rw = rh or H, rw or W
    rfw = torch.randn(B, 2 * K, rh, rw, generator=g) * 5.0
    rbw = torch.randn(B, 2 * K, rh, rw, generator=g) * 5.0
    return masks, fw, bw, rfw, rbw


def run_case(ref, shape, kwargs, seed):
    K, H, W = shape["K"], shape["H"], shape["W"]
    inputs32 = make_inputs(seed=seed, **shape)
    out = {}
    kw = dict(mask_layer=K, mask_size=(H, W), **kwargs)
    head32 = ref_loader.build_reference_head(ref, torch.float32, seed=1, **kw)
    params = {k: v.detach().clone() for k, v in head32.state_dict().items()}
    for tag, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        head = ref_loader.build_reference_head(ref, dtype, seed=1, **kw)
        head.load_state_dict({k: v.to(dtype) for k, v in params.items()})
        ins = [t.to(dtype) for t in inputs32]
        flows, loss, grads = ref_loader.run_reference(head, *ins, gbar=0.7)
        for k, v in loss.items():
            out[f"{tag}.loss.{k}"] = np.float64(v)
        for k, v in flows.items():
            if len(v):
                out[f"{tag}.flows.{k}"] = v[0].detach().numpy()
        for k in ("d_masks", "d_resid_fw", "d_resid_bw"):
            out[f"{tag}.{k}"] = grads[k].numpy()
        for k, v in grads["params"].items():
            out[f"{tag}.dparam.{k}"] = v.numpy()
    for name, t in zip(("masks", "fw", "bw", "rfw", "rbw"), inputs32):
        out[f"in.{name}"] = t.numpy()
    for k, v in params.items():
        out[f"param.{k}"] = v.numpy()
    out["meta"] = np.array(json.dumps(dict(shape=shape, kwargs=kwargs, seed=seed, gbar=0.7)))
    return out


def checksum(t: torch.Tensor):
    t = t.double()
    idx = torch.linspace(0, t.numel() - 1, 16).long()
    return dict(sum=float(t.sum()), l2=float(t.norm()), abs_sum=float(t.abs().sum()),
                sample=[float(x) for x in t.flatten()[idx]])


def full_size_scalars(ref):
    """Reference fp32 at C1 (B=2,K=4,480x854, Cf=64) -- minutes of CPU; loss-core proxy head (Cf=2,k=1) too."""
    res = {}
    for name, kw in (("c1_free_l1_full_head", dict(free_residual=True, clamp_flow_t=20.0)),
                     ("c1_free_l1_proxy_head", dict(free_residual=True, clamp_flow_t=20.0,
                                                    num_flow_feat_channels=2, flow_feat_before_agg_kernel_size=1)),
                     ("c1_affine_l1_proxy_head", dict(free_residual_with_affine=True, clamp_flow_t=20.0,
                                                      num_flow_feat_channels=2, flow_feat_before_agg_kernel_size=1))):
        B, K, H, W = 2, 4, 480, 854
        ins = make_inputs(B, K, H, W, seed=0)
        head = ref_loader.build_reference_head(ref, torch.float32, seed=1, mask_layer=K, mask_size=(H, W), **kw)
        flows, loss, grads = ref_loader.run_reference(head, *ins, gbar=1.0)
        res[name] = dict(shape=dict(B=B, K=K, H=H, W=W), kwargs=kw, seed=0, param_seed=1, loss=loss,
                         d_masks=checksum(grads["d_masks"]), d_resid_fw=checksum(grads["d_resid_fw"]),
                         d_resid_bw=checksum(grads["d_resid_bw"]),
                         dparams={k: checksum(v) for k, v in grads["params"].items()},
                         pred_flow=checksum(flows["pred_flow"][0]))
        print(name, loss, flush=True)
    return res


def main():
    assert ref_loader.reference_available(), "needs /root/reference"
    ref = ref_loader.load_reference_module()
    for i, (name, (shape, kwargs)) in enumerate(CASES.items()):
        out = run_case(ref, shape, kwargs, seed=100 + i)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: float(v) for k, v in out.items() if k.startswith("f64.loss")}, flush=True)
    if "--no-full" not in sys.argv:
        with open(os.path.join(HERE, "full_size_scalars.json"), "w") as f:
            json.dump(full_size_scalars(ref), f, indent=1)


if __name__ == "__main__":
    main()
