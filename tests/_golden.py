"""Helpers shared by the oracle and GPU parity tests: load the fixtures that
oracle/gen_golden.py produced by running the reference."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
CBAM_CASES = ("basic", "beta", "edge", "nomask", "rawmask", "mask3d", "gate_eval", "gate_det", "k5")
PARAM_KEYS = ("cam_mlp.0.weight", "cam_mlp.0.bias", "cam_mlp.2.weight", "cam_mlp.2.bias", "sam_conv.weight", "beta")


def load_cbam(tag: str):
    z = np.load(GOLDEN / f"cbam_{tag}.npz")
    rec = {k: z[k] for k in z.files}
    C, r, k, use_sig, gated = (int(v) for v in rec["cfg"])
    rec["C"], rec["r"], rec["k"], rec["use_sigmoid_mask"], rec["gated"] = C, r, k, bool(use_sig), bool(gated)
    return rec


def t(a, dtype=None):
    out = torch.from_numpy(np.ascontiguousarray(a))
    return out if dtype is None else out.to(dtype)


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """max|got-ref| / max|ref|  (the SURVEY section 8d gate)."""
    ref = ref.double()
    den = ref.abs().max().item()
    return (got.double() - ref).abs().max().item() / (den if den > 0 else 1.0)


# ---- full-size YOLOv8n neck shapes at batch 2: inputs are regenerated from seeds (not stored), the reference's outputs are
# stored sampled (oracle/gen_golden_large.py).  These shapes run multi-CTA clusters in both directions (DSMEM + halo rows).
LARGE_CASES = {"p3": (2, 64, 80, 80), "p4": (2, 128, 40, 40), "p5": (2, 256, 20, 20)}
LARGE_SAMPLES = 4096


def large_inputs(tag: str):
    """(x, mask, g, beta, module seed, sample index) of a large case -- the same code runs in the generator and in the tests."""
    B, C, H, W = LARGE_CASES[tag]
    gen = torch.Generator().manual_seed(1000 + C)
    x = torch.randn(B, C, H, W, generator=gen)
    mask = torch.randn(B, 1, H, W, generator=gen) * 2.0
    mask[1, :, : H // 4] = -9.0  # a band that is masked out
    g = torch.randn(B, C, H, W, generator=gen)
    idx = torch.randperm(B * C * H * W, generator=gen)[:LARGE_SAMPLES]
    return x, mask, g, 0.35, C, idx
