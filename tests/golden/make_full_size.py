"""Reference (UNMODIFIED head, fp32, CPU) loss scalars + gradient checksums at the remaining BASELINE.json shapes:
C4 (K = 2 / 8 at 480x854), C5 (240x427, 1080x1920) and the affine configs at full size (K = 3 FBMS, K = 8).  The
loss-core proxy head (Cf = 2, k = 1) and B = 1 keep the CPU time in minutes.  Build container only:

    python tests/golden/make_full_size.py        # writes full_size_scalars_r2.json
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from make_golden import checksum, make_inputs  # noqa: E402

PROXY = dict(num_flow_feat_channels=2, flow_feat_before_agg_kernel_size=1, clamp_flow_t=20.0)
CASES = {
    "c4_k2_free": (dict(B=1, K=2, H=480, W=854), dict(free_residual=True, **PROXY)),
    "c4_k8_free": (dict(B=1, K=8, H=480, W=854), dict(free_residual=True, **PROXY)),
    "c5_240x427_free": (dict(B=1, K=4, H=240, W=427), dict(free_residual=True, **PROXY)),
    "c5_1080x1920_free": (dict(B=1, K=4, H=1080, W=1920), dict(free_residual=True, **PROXY)),
    "fbms_k3_affine": (dict(B=1, K=3, H=480, W=854), dict(free_residual_with_affine=True, **PROXY)),
    "c4_k8_affine": (dict(B=1, K=8, H=480, W=854), dict(free_residual_with_affine=True, **PROXY)),
}


def main():
    assert ref_loader.reference_available(), "needs /root/reference"
    ref = ref_loader.load_reference_module()
    torch.set_num_threads(os.cpu_count())
    res = {}
    for name, (shape, kw) in CASES.items():
        B, K, H, W = shape["B"], shape["K"], shape["H"], shape["W"]
        ins = make_inputs(B, K, H, W, seed=0)
        head = ref_loader.build_reference_head(ref, torch.float32, seed=1, mask_layer=K, mask_size=(H, W), **kw)
        flows, loss, grads = ref_loader.run_reference(head, *ins, gbar=1.0)
        res[name] = dict(shape=shape, kwargs=kw, seed=0, param_seed=1, loss=loss,
                         d_masks=checksum(grads["d_masks"]), d_resid_fw=checksum(grads["d_resid_fw"]),
                         d_resid_bw=checksum(grads["d_resid_bw"]),
                         dparams={k: checksum(v) for k, v in grads["params"].items()},
                         pred_flow=checksum(flows["pred_flow"][0]))
        print(name, loss, flush=True)
        with open(os.path.join(HERE, "full_size_scalars_r2.json"), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
