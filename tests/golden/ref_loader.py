"""Load the UNMODIFIED reference head from /root/reference (build container only).

Used by ``make_golden.py`` (to generate the committed fixtures) and by the optional
``tests/test_reference_live.py`` (skipped when /root/reference is absent, e.g. on the GPU box).
Recipe from SURVEY.md 8(c): the head only needs ``utils.get_logger``; the real ``utils`` package
pulls in mmseg, which is not installed.
"""
import contextlib
import importlib.util
import logging
import os
import sys
import types

import torch

REF_FILE = "/root/reference/models/flow_aggregation_head_with_residual.py"


def reference_available() -> bool:
    return os.path.exists(REF_FILE)


def load_reference_module():
    saved = sys.modules.get("utils")
    stub = types.ModuleType("utils")
    stub.get_logger = lambda: logging.getLogger("rcf_reference")
    sys.modules["utils"] = stub
    try:
        spec = importlib.util.spec_from_file_location("rcf_reference_head", REF_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules["utils"] = saved
        else:
            del sys.modules["utils"]
    return mod


@contextlib.contextmanager
def cuda_is_identity():
    """The reference hard-codes ``coord_map.cuda()`` (:143,:146); make it a no-op on CPU boxes."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


@contextlib.contextmanager
def float_is_identity():
    """For the fp64 arbiter: neutralise the forced ``.float()`` around linalg.solve (:216-217)."""
    orig = torch.Tensor.float
    torch.Tensor.float = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.float = orig


def build_reference_head(ref_mod, dtype=torch.float32, seed=1, **kwargs):
    torch.manual_seed(seed)
    with cuda_is_identity():
        head = ref_mod.FlowAggregationHeadWithResidual(args=None, create_flownet=True, **kwargs)
    if dtype == torch.float64:
        head = head.double()
        if hasattr(head, "coord_map"):
            head.coord_map = head.coord_map.double()
    return head


def run_reference(head, masks, fw, bw, rfw, rbw, gbar=1.0):
    """fwd + bwd; returns (flows, loss dict of floats, grads dict).  Tensors must be leaf tensors."""
    masks = masks.clone().requires_grad_(True)
    rfw = rfw.clone().requires_grad_(True)
    rbw = rbw.clone().requires_grad_(True)
    imgs = torch.zeros(masks.shape[0], 2, 3, 8, 8, dtype=masks.dtype)
    for p in head.parameters():
        p.grad = None
    ctx = float_is_identity() if masks.dtype == torch.float64 else contextlib.nullcontext()
    with ctx:
        flows, loss = head(imgs, masks, fw.clone(), bw.clone(), rfw, rbw)
        (loss["seg"] * gbar).backward()
    grads = {"d_masks": masks.grad, "d_resid_fw": rfw.grad, "d_resid_bw": rbw.grad,
             "params": {k: p.grad for k, p in head.named_parameters()}}
    return flows, {k: float(v) for k, v in loss.items()}, grads
