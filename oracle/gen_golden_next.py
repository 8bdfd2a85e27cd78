"""Generate the fixtures of the "next" rows by RUNNING THE REFERENCE (authoring container only):

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 YOLO_CONFIG_DIR=/tmp/ulcfg python oracle/gen_golden_next.py

  tests/golden/eca_*.npz   mga_yolo.nn.modules.masked_eca.MaskECA            forward + torch autograd backward (fp32 and fp64)
  tests/golden/head_*.npz  mga_yolo.nn.modules.segmentation.MGAMaskHead      `head` 3x3 conv tail: forward + autograd (fp32 and fp64)
  tests/golden/gate_*.npz  mga_yolo.nn.modules.probmaskgater.ProbMaskGater   train-mode outputs with the uniform noise captured
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


def eca_case(tag, B, C, H, W, *, seed, beta=0.0, mask_kind="logits", use_sigmoid_mask=True):
    from mga_yolo.nn.modules.masked_eca import MaskECA

    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=gen)
    g = torch.randn(B, C, H, W, generator=gen)
    mask = None
    if mask_kind == "logits":
        mask = torch.randn(B, 1, H, W, generator=gen) * 2
    elif mask_kind == "edge":  # one sample with a tiny mask (GAP fall-back), one 3-D mask batch
        mask = torch.randn(B, 1, H, W, generator=gen)
        mask[0] = -20.0
    elif mask_kind == "raw3d":
        mask = (torch.rand(B, H, W, generator=gen) > 0.6).float()
    torch.manual_seed(seed)
    mod = MaskECA(C, use_sigmoid_mask=use_sigmoid_mask)
    with torch.no_grad():
        mod.beta.fill_(beta)
    rec = {"x": x.numpy(), "g": g.numpy(), "has_mask": np.array(mask is not None), "w1d": mod.conv1d.weight.detach().numpy().copy(),
           "beta": np.array(beta, dtype=np.float32), "cfg": np.array([C, int(use_sigmoid_mask)], dtype=np.int64)}
    if mask is not None:
        rec["mask"] = mask.numpy()
    for suffix, dt in (("", torch.float32), ("_f64", torch.float64)):
        m = mod.double() if dt == torch.float64 else mod.float()
        xi = x.to(dt).clone().requires_grad_(True)
        mi = None if mask is None else mask.to(dt).clone().requires_grad_(True)
        for p in m.parameters():
            p.grad = None
        out = m(xi if mi is None else [xi, mi])
        out.backward(g.to(dt))
        rec["out" + suffix] = out.detach().numpy()
        rec["dx" + suffix] = xi.grad.numpy()
        if mi is not None:
            rec["dmask" + suffix] = mi.grad.numpy()
        rec["d.conv1d.weight" + suffix] = m.conv1d.weight.grad.numpy().copy()
        rec["d.beta" + suffix] = m.beta.grad.numpy().copy()
    mod.float()
    np.savez_compressed(OUT / f"eca_{tag}.npz", **rec)
    print("eca", tag, rec["out"].shape, "k =", mod.conv1d.weight.shape[-1])


def head_case(tag, B, Cin, hidden, H, W, *, seed):
    from mga_yolo.nn.modules.segmentation import MGAMaskHead

    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    head = MGAMaskHead(Cin, hidden)
    with torch.no_grad():
        head.head.bias.fill_(0.1)
    feat = torch.randn(B, hidden, H, W, generator=gen)   # the tail's input: output of head.proj
    g = torch.randn(B, 1, H, W, generator=gen)
    rec = {"feat": feat.numpy(), "g": g.numpy(), "w": head.head.weight.detach().numpy().copy(), "b": head.head.bias.detach().numpy().copy()}
    for suffix, dt in (("", torch.float32), ("_f64", torch.float64)):
        conv = head.head.double() if dt == torch.float64 else head.head.float()
        fi = feat.to(dt).clone().requires_grad_(True)
        conv.weight.grad = None
        conv.bias.grad = None
        out = conv(fi)
        out.backward(g.to(dt))
        rec["out" + suffix] = out.detach().numpy()
        rec["dfeat" + suffix] = fi.grad.numpy()
        rec["dw" + suffix] = conv.weight.grad.numpy().copy()
        rec["db" + suffix] = conv.bias.grad.numpy().copy()
    head.head.float()
    np.savez_compressed(OUT / f"head_{tag}.npz", **rec)
    print("head", tag, rec["out"].shape)


def gate_case(tag, mode, *, seed, tau=0.5, p_min=0.0):
    """Train-mode ProbMaskGater (probmaskgater.py:73-95) with the uniform noise it drew captured (torch.rand is wrapped for the call), so
    that a kernel fed the same uniforms must reproduce the output."""
    from mga_yolo.nn.modules import probmaskgater as pg

    gen = torch.Generator().manual_seed(seed)
    p_in = torch.rand(2, 1, 12, 20, generator=gen) * 1.4 - 0.2   # some values outside [0,1]: the clamp matters
    gater = pg.ProbMaskGater(mode=mode, tau=tau, p_min=p_min, seed=None)
    gater.train()
    drawn = []
    real = torch.rand

    def spy(*a, **k):
        u = real(*a, **k)
        drawn.append(u.clone())
        return u

    torch.manual_seed(seed)
    torch.rand = spy
    try:
        pi = p_in.clone().requires_grad_(True)
        out = gater(pi)
    finally:
        torch.rand = real
    gup = torch.randn(out.shape, generator=gen)
    out.backward(gup)
    rec = {"p": p_in.numpy(), "out": out.detach().numpy(), "g": gup.numpy(), "dp": pi.grad.numpy(),
           "cfg": np.array([tau, p_min, gater.threshold], dtype=np.float64), "mode": np.array(mode)}
    for i, u in enumerate(drawn):
        rec[f"u{i}"] = u.numpy()   # (before the reference's clamp_(1e-6, 1 - 1e-6), which happens in place on the returned tensor)
    rec["n_u"] = np.array(len(drawn))
    np.savez_compressed(OUT / f"gate_{tag}.npz", **rec)
    print("gate", tag, mode, "uniform draws:", len(drawn))


if __name__ == "__main__":
    eca_case("basic", 3, 64, 12, 16, seed=1)
    eca_case("beta", 2, 128, 10, 10, seed=2, beta=0.4)
    eca_case("edge", 3, 32, 9, 7, seed=3, beta=-0.3, mask_kind="edge")
    eca_case("nomask", 2, 48, 8, 8, seed=4, mask_kind="none")
    eca_case("raw3d", 2, 256, 6, 10, seed=5, beta=0.2, mask_kind="raw3d", use_sigmoid_mask=False)
    head_case("p3", 2, 64, 32, 20, 24, seed=11)
    head_case("odd", 1, 128, 64, 7, 9, seed=12)
    gate_case("gumbel", "gumbel", seed=21, tau=0.5)
    gate_case("gumbel_pmin", "gumbel", seed=22, tau=1.0, p_min=0.2)
    gate_case("hard_st", "hard_st", seed=23, tau=0.7)
