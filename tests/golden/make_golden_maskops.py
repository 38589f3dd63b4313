"""Generate tests/golden/aux/mask_losses.npz from the UNMODIFIED reference code (build container only).

models/compactness_head.py imports nothing but torch and is loaded as a module.  models/rcf_model.py cannot be imported
here (mmseg / pytorch_lightning are not installed), so the bodies of RCFModel.get_entropy_loss / get_pl_loss / get_crf_loss
are cut out of the unmodified file with `ast` and executed as they are, with a namespace standing in for `self`; the two
lines that produce their inputs (:433-434, softmax and log_softmax OF the softmax) are quoted in this script.
"""
import ast
import importlib.util
import os
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/models"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def extract_methods(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    lu = load("/root/reference/utils/loss_utils.py", "ref_loss_utils_for_sharpen")
    ns = {"torch": torch, "F": F, "utils": types.SimpleNamespace(sharpen=lu.sharpen)}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = ast.get_source_segment(src, node)
            lines = code.split("\n")
            indent = len(lines[0]) - len(lines[0].lstrip()) if lines[0].startswith(" ") else 0
            exec("\n".join(l[indent:] if l.startswith(" " * indent) else l for l in lines), ns)      # noqa: S102
    return ns


def main():
    ch = load(os.path.join(REF, "compactness_head.py"), "ref_compactness_head")
    ns = extract_methods(os.path.join(REF, "rcf_model.py"), {"get_entropy_loss", "get_pl_loss", "get_crf_loss", "get_sharpen_loss"})
    out = {}
    cases = [dict(name="k4", shape=(2, 2, 4, 12, 16), compact=0, oc=2, th=-1.0, wp=1.0, wn=1.0),
             dict(name="k3_th", shape=(2, 2, 3, 9, 7), compact=1, oc=0, th=0.5, wp=2.0, wn=0.5),
             dict(name="k5", shape=(1, 2, 5, 6, 10), compact=4, oc=4, th=-1.0, wp=0.3, wn=1.7)]
    g = torch.Generator().manual_seed(11)
    for c in cases:
        B, I, K, H, W = c["shape"]
        logits = (torch.randn(B, I, K, H, W, generator=g, dtype=torch.float64) * 3).requires_grad_(True)
        w_mask = torch.randn(B, I, K, H, W, generator=g, dtype=torch.float64)
        pl = torch.rand(B, I, H, W, generator=g, dtype=torch.float64)
        all_pred_mask = F.softmax(logits, dim=2)                          # rcf_model.py:433
        log_all_pred_mask = F.log_softmax(all_pred_mask, dim=2)           # rcf_model.py:434
        self_pl = types.SimpleNamespace(pl_mask_pos_th=c["th"], pl_pos_weight=c["wp"], pl_neg_weight=c["wn"],
                                        args=types.SimpleNamespace(object_channel=c["oc"]))
        self_crf = types.SimpleNamespace(crf_mask_pos_th=c["th"], crf_pos_weight=c["wp"], crf_neg_weight=c["wn"],
                                         args=types.SimpleNamespace(object_channel=c["oc"]))
        ent = ns["get_entropy_loss"](None, all_pred_mask, log_all_pred_mask)
        pl_loss = ns["get_pl_loss"](self_pl, all_pred_mask, pl)
        crf_loss = ns["get_crf_loss"](self_crf, all_pred_mask, pl)
        assert torch.equal(pl_loss, crf_loss)                             # the two reference methods are the same formula
        head = ch.CompactnessHead(args=types.SimpleNamespace(object_channel=None), compact_channel=c["compact"])
        comp = head.get_compactness_loss(all_pred_mask)
        coef = (0.7, 1.3, 2.1)
        total = (all_pred_mask * w_mask).sum() + coef[0] * ent + coef[1] * comp + coef[2] * pl_loss
        (gl,) = torch.autograd.grad(total, logits, retain_graph=True)
        # get_sharpen_loss (:350-374), both live variants, each differentiated on its own
        for tag, oa in (("kl", False), ("object_hinge", True)):
            self_sh = types.SimpleNamespace(object_aware_sharpening=oa, t_sharpen=0.25)
            sh = ns["get_sharpen_loss"](self_sh, all_pred_mask, log_all_pred_mask, object_channel=c["oc"])
            (gsh,) = torch.autograd.grad(1.7 * sh, logits, retain_graph=True)
            out[f"{c['name']}.sharpen.{tag}"] = np.array(float(sh))
            out[f"{c['name']}.sharpen.{tag}.dlogits"] = gsh.numpy()
        n = c["name"]
        out[f"{n}.logits"], out[f"{n}.w_mask"], out[f"{n}.pl"] = logits.detach().numpy(), w_mask.numpy(), pl.numpy()
        out[f"{n}.masks"] = all_pred_mask.detach().numpy()
        out[f"{n}.losses"] = np.array([float(ent), float(comp), float(pl_loss)])
        out[f"{n}.dlogits"] = gl.numpy()
        out[f"{n}.cfg"] = np.array([c["compact"], c["oc"], c["th"], c["wp"], c["wn"], *coef])
    np.savez_compressed(os.path.join(HERE, "aux", "mask_losses.npz"), **out)
    print("written", sorted(out))


if __name__ == "__main__":
    main()
