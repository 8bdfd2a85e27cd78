"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (authoring container only).

    PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 YOLO_CONFIG_DIR=/tmp/ulcfg \
        python oracle/gen_golden.py

/root/reference does not exist on the GPU box, so the outputs are committed as small
fixtures; nothing under tests/, smoke() or bench.py reads the reference at run time.

What is called:
  mga_yolo.nn.modules.masked_cbam.MaskCBAM            (forward + torch autograd backward)
  mga_yolo.utils.mask_utils.MaskUtils.downsample_mask / downsample_mask_prob (-> cv2)
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
OUT.mkdir(parents=True, exist_ok=True)


def _randn(gen, *shape, scale=1.0):
    return torch.randn(*shape, generator=gen, dtype=torch.float64).mul_(scale).float()


def run_ref(tag, x, mask, g, *, C, r=16, k=7, use_sigmoid_mask=True, beta=0.0, seed=0, prob_env=None, train=True):
    """Run the reference block in fp32 and in fp64 on identical inputs; dump everything."""
    for key in ("MGA_PROB_MODE", "MGA_PROB_APPROACH"):
        os.environ.pop(key, None)
    if prob_env:
        os.environ.update(prob_env)
    from mga_yolo.nn.modules.masked_cbam import MaskCBAM

    torch.manual_seed(seed)
    mod = MaskCBAM(C, r=r, spatial_k=k, use_sigmoid_mask=use_sigmoid_mask)
    with torch.no_grad():
        mod.beta.fill_(beta)
    mod.train(train)
    rec = {"x": x.numpy(), "g": g.numpy(), "has_mask": np.array(mask is not None)}
    if mask is not None:
        rec["mask"] = mask.numpy()
    for name, t in mod.state_dict().items():
        rec["p." + name] = t.detach().numpy().copy()
    rec["cfg"] = np.array([C, r, mod.k, int(use_sigmoid_mask), int(bool(prob_env))], dtype=np.int64)

    for suffix, dt in (("", torch.float32), ("_f64", torch.float64)):
        m = mod.double() if dt == torch.float64 else mod.float()
        xi = x.to(dt).clone().requires_grad_(True)
        mi = None if mask is None else mask.to(dt).clone().requires_grad_(True)
        for prm in m.parameters():
            prm.grad = None
        out = m(xi if mi is None else [xi, mi])
        out.backward(g.to(dt))
        rec["out" + suffix] = out.detach().numpy()
        rec["dx" + suffix] = xi.grad.numpy()
        if mi is not None:
            rec["dmask" + suffix] = mi.grad.numpy()
        for name, prm in m.named_parameters():
            rec["d." + name + suffix] = prm.grad.detach().numpy().copy()
    np.savez_compressed(OUT / f"cbam_{tag}.npz", **rec)
    for key in ("MGA_PROB_MODE", "MGA_PROB_APPROACH"):
        os.environ.pop(key, None)
    print("wrote", tag, {k_: v.shape for k_, v in rec.items() if k_ in ("x", "mask")})


def gen_cbam():
    gen = torch.Generator().manual_seed(1234)
    # 1. plain case, vector-friendly plane (S = 120)
    B, C, H, W = 3, 32, 10, 12
    run_ref("basic", _randn(gen, B, C, H, W), _randn(gen, B, 1, H, W), _randn(gen, B, C, H, W), C=C)
    # 2. non-zero beta, wider logits, h = 3
    B, C, H, W = 2, 48, 8, 8
    run_ref("beta", _randn(gen, B, C, H, W), _randn(gen, B, 1, H, W, scale=2.0), _randn(gen, B, C, H, W), C=C, beta=0.3, seed=1)
    # 3. fall-back branches + ragged plane (S = 63): sample 1 is all logits -20 (tiny mask AND no
    #    valid pixel), sample 2 is all logits -1 (mask mean fine, no pixel above 0.5), sample 3
    #    is a hard +-4 mask (masked_cbam.py:97-102,118-121)
    B, C, H, W = 4, 16, 9, 7
    mk = _randn(gen, B, 1, H, W)
    mk[1] = -20.0
    mk[2] = -1.0
    mk[3] = torch.where(mk[3] > 0, torch.tensor(4.0), torch.tensor(-4.0))
    run_ref("edge", _randn(gen, B, C, H, W), mk, _randn(gen, B, C, H, W), C=C, r=4, beta=-0.2, seed=2)
    # 4. no mask -> vanilla CBAM with a zero third plane (masked_cbam.py:90-91,107-108,137-138)
    B, C, H, W = 2, 32, 8, 12
    run_ref("nomask", _randn(gen, B, C, H, W), None, _randn(gen, B, C, H, W), C=C, seed=3)
    # 5. raw {0,1} mask, use_sigmoid_mask=False; sample 1 is an all-zero mask (sum < eps)
    B, C, H, W = 3, 32, 8, 8
    mk = (torch.rand(B, 1, H, W, generator=gen) > 0.6).float()
    mk[1] = 0.0
    run_ref("rawmask", _randn(gen, B, C, H, W), mk, _randn(gen, B, C, H, W), C=C, use_sigmoid_mask=False, beta=0.1, seed=4)
    # 6. (B,H,W) mask
    B, C, H, W = 2, 16, 6, 8
    run_ref("mask3d", _randn(gen, B, C, H, W), _randn(gen, B, H, W), _randn(gen, B, C, H, W), C=C, r=8, seed=5)
    # 7. ProbMaskGater on the deterministic path (probmaskgater.py:77,82-83): logits clamped to [0,1]
    B, C, H, W = 2, 32, 8, 8
    run_ref("gate_eval", _randn(gen, B, C, H, W), _randn(gen, B, 1, H, W), _randn(gen, B, C, H, W), C=C, seed=6,
            prob_env={"MGA_PROB_MODE": "1", "MGA_PROB_APPROACH": "gumbel"}, train=False)
    run_ref("gate_det", _randn(gen, B, C, H, W), _randn(gen, B, 1, H, W), _randn(gen, B, C, H, W), C=C, seed=7,
            prob_env={"MGA_PROB_MODE": "1", "MGA_PROB_APPROACH": "deterministic"}, train=True)
    # 8. 5x5 spatial kernel, r = 2
    B, C, H, W = 2, 8, 12, 12
    run_ref("k5", _randn(gen, B, C, H, W), _randn(gen, B, 1, H, W), _randn(gen, B, C, H, W), C=C, r=2, k=5, seed=8)


def synth_mask(rng, h, w):
    """Binary mask with solid blobs, thin lines and a 50% noise patch (so block counts land
    on both sides of s*s/2, the INTER_AREA rounding boundary)."""
    m = np.zeros((h, w), np.uint8)
    for _ in range(4):
        y0, x0 = rng.integers(0, h), rng.integers(0, w)
        hh, ww = rng.integers(2, max(3, h // 3)), rng.integers(2, max(3, w // 3))
        m[y0 : y0 + hh, x0 : x0 + ww] = 1
    for _ in range(3):
        y = rng.integers(0, h)
        m[y, rng.integers(0, w // 2) : rng.integers(w // 2, w)] = 1
    y0, x0 = h // 3, w // 3
    patch = m[y0 : y0 + h // 3, x0 : x0 + w // 3]
    patch[...] = rng.random(patch.shape) > 0.5
    # exact half-filled blocks at a few strides (rounding ties)
    m[:8, :8] = 0
    m[:4, :8] = 1
    if h >= 32 and w >= 64:
        m[16:32, 32:48] = 0
        m[16:24, 32:48] = 1
    return m


def gen_masks():
    from mga_yolo.utils.mask_utils import MaskUtils

    rng = np.random.default_rng(7)
    rec = {}
    sizes = [(64, 64), (96, 128), (160, 96), (100, 75), (33, 50), (8, 8), (5, 3)]
    strides = (4, 8, 16, 32)
    rec["sizes"] = np.array(sizes)
    rec["strides"] = np.array(strides)
    for si, (h, w) in enumerate(sizes):
        m = synth_mask(rng, h, w) if min(h, w) >= 8 else (rng.random((h, w)) > 0.5).astype(np.uint8)
        rec[f"src{si}"] = m
        for s in strides:
            for method, bridge in (("nearest", "1"), ("area", "1"), ("area", "0"), ("maxpool", "1"),
                                   ("skeleton_bresenham", "1"), ("skeleton_bresenham", "0")):
                os.environ["MGA_MASK_METHOD"] = method
                os.environ["MGA_MASK_BRIDGE"] = bridge
                os.environ.pop("MGA_MASK_THRESH", None)
                rec[f"bin{si}_{s}_{method}_{bridge}"] = MaskUtils.downsample_mask(m, s)
            for method in ("avgpool", "nearest", "area"):
                rec[f"prob{si}_{s}_{method}"] = MaskUtils.downsample_mask_prob(m, s, method=method)
    for key in ("MGA_MASK_METHOD", "MGA_MASK_BRIDGE"):
        os.environ.pop(key, None)
    np.savez_compressed(OUT / "mask_downsample.npz", **rec)
    print("wrote mask goldens:", len(rec), "arrays")


if __name__ == "__main__":
    gen_cbam()
    gen_masks()
