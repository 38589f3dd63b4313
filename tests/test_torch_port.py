"""Pin oracle/torch_port.py (the CPU baseline that bench.py times) to the reference's fp32 outputs."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle.torch_port import PortedHead


def test_port_matches_reference_fp32(golden):
    kw = {k: v for k, v in golden.head_kwargs().items()}
    head = PortedHead(**kw)
    head.load_state_dict({k: torch.from_numpy(v) for k, v in golden.params.items()})
    masks, fw, bw, rfw, rbw = [torch.from_numpy(a.copy()) for a in golden.inputs]
    masks.requires_grad_(True); rfw.requires_grad_(True); rbw.requires_grad_(True)
    flows, loss = head(torch.zeros(masks.shape[0], 2, 3, 8, 8), masks, fw, bw, rfw, rbw)
    (loss["seg"] * golden.gbar).backward()
    assert float(loss["seg"]) == pytest.approx(float(golden.ref("f32", "loss.seg")), rel=2e-6)
    quad = golden.kwargs.get("free_residual_with_affine_quadratic", False)
    tol = 5e-2 if quad else 2e-4       # fp32 vs fp32 of an ill-conditioned solve (reference noise floor)
    assert rel_l2(masks.grad.numpy(), golden.ref("f32", "d_masks")) < tol
    assert rel_l2(rfw.grad.numpy(), golden.ref("f32", "d_resid_fw")) < tol
    for k, v in flows.items():
        if v:
            assert rel_l2(v[0].detach().numpy(), golden.ref("f32", "flows." + k)) < 1e-5
