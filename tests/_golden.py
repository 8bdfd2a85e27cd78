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
