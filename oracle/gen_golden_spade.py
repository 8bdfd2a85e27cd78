"""Generate the MaskSPADE fixtures by RUNNING THE REFERENCE (authoring container only):

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 YOLO_CONFIG_DIR=/tmp/ulcfg python oracle/gen_golden_spade.py

  tests/golden/spade_*.npz   mga_yolo.nn.modules.masked_spade.MaskSPADE   forward + torch autograd backward (fp32 and fp64),
                             with the block's own gamma / beta captured (the operands of the CUDA feature-side op)
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
sys.dont_write_bytecode = True
if "/root/reference" not in sys.path:
    sys.path.insert(0, "/root/reference")

import numpy as np
import torch

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
KEYS = ("shared.0.weight", "shared.0.bias", "conv_gamma.weight", "conv_gamma.bias", "conv_beta.weight", "conv_beta.bias")


def spade_case(tag, B, C, H, W, *, seed, hidden=16, mask_kind="logits", use_sigmoid_mask=True, mask_hw=None):
    from mga_yolo.nn.modules.masked_spade import MaskSPADE

    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=gen) * 1.5 + 0.3
    g = torch.randn(B, C, H, W, generator=gen)
    mh, mw = mask_hw or (H, W)
    mask = None
    if mask_kind == "logits":
        mask = torch.randn(B, 1, mh, mw, generator=gen) * 2
    elif mask_kind == "raw3d":
        mask = (torch.rand(B, mh, mw, generator=gen) > 0.6).float()
    torch.manual_seed(seed)
    mod = MaskSPADE(C, hidden=hidden, use_sigmoid_mask=use_sigmoid_mask)
    with torch.no_grad():  # the reference initialises the biases to zero: move them so that their gradients are exercised
        for k, p in mod.named_parameters():
            if k.endswith("bias"):
                p.normal_(0.0, 0.2, generator=gen)
    rec = {"x": x.numpy(), "g": g.numpy(), "has_mask": np.array(mask is not None),
           "cfg": np.array([C, hidden, int(use_sigmoid_mask)], dtype=np.int64)}
    for k in KEYS:
        rec["p." + k] = mod.state_dict()[k].numpy().copy()
    if mask is not None:
        rec["mask"] = mask.numpy()
    for suffix, dt in (("", torch.float32), ("_f64", torch.float64)):
        m = mod.double() if dt == torch.float64 else mod.float()
        xi = x.to(dt).clone().requires_grad_(True)
        mi = None if mask is None else mask.to(dt).clone().requires_grad_(True)
        for p in m.parameters():
            p.grad = None
        cap = {}
        hooks = [m.conv_gamma.register_forward_hook(lambda _m, _i, o: cap.__setitem__("gamma", o.detach().clone())),
                 m.conv_beta.register_forward_hook(lambda _m, _i, o: cap.__setitem__("beta", o.detach().clone()))]
        out = m(xi if mi is None else [xi, mi])
        out.backward(g.to(dt))
        for h in hooks:
            h.remove()
        rec["out" + suffix] = out.detach().numpy()
        rec["dx" + suffix] = xi.grad.numpy()
        if mi is not None:
            rec["dmask" + suffix] = mi.grad.numpy()
            rec["gamma" + suffix] = cap["gamma"].numpy()
            rec["beta" + suffix] = cap["beta"].numpy()
            for k in KEYS:
                rec["d." + k + suffix] = dict(m.named_parameters())[k].grad.numpy().copy()
    mod.float()
    np.savez_compressed(OUT / f"spade_{tag}.npz", **rec)
    print("spade", tag, rec["out"].shape)


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    spade_case("basic", 2, 16, 12, 20, seed=11)
    spade_case("nomask", 2, 8, 10, 12, seed=12, mask_kind="none")
    spade_case("raw3d", 2, 8, 16, 16, seed=13, mask_kind="raw3d", use_sigmoid_mask=False)
    spade_case("odd", 3, 5, 7, 9, seed=14, hidden=8)          # H*W not a multiple of the vector width: scalar path
    spade_case("resize", 2, 8, 12, 12, seed=15, mask_hw=(24, 24))  # bilinear resize of the mask to the feature size
    spade_case("p5", 1, 32, 20, 20, seed=16)                   # 400-pixel rows: small CTAs
