"""CPU restatements for the SURVEY.md section 8f rows next to the block (TEST INFRASTRUCTURE ONLY, never imported by mga_yolo_b200/).

  head_tail_*   MGAMaskHead.head: Conv2d(hidden, 1, 3, padding=1, bias=True)      /root/reference/mga_yolo/nn/modules/segmentation.py:94,107-110
                (explicit shifted sums, closed-form backward: no F.conv2d, no autograd)
  gate_*        ProbMaskGater.forward in train mode given the two uniform draws    /root/reference/mga_yolo/nn/modules/probmaskgater.py:59-95
  philox4x32    the counter-based generator the CUDA gate kernel draws its uniforms from (Salmon et al., SC'11; 10 rounds), for the
                known-answer test and the bit-level check of the kernel's noise
  collate_masks zero-pad + stack of per-sample pyramid masks                        /root/reference/mga_yolo/data/dataset.py:149-169
Pinned against fixtures made by running the reference (oracle/gen_golden_next.py) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch


# ------------------------------------------------------------------ MGAMaskHead tail (3x3 conv, one output channel)
def _shift(t, dy, dx):
    """t (B,C,H,W) -> u with u[..., y, x] = t[..., y + dy, x + dx], zero outside."""
    B, C, H, W = t.shape
    out = torch.zeros_like(t)
    ys, ye = max(0, -dy), min(H, H - dy)
    xs, xe = max(0, -dx), min(W, W - dx)
    if ys < ye and xs < xe:
        out[:, :, ys:ye, xs:xe] = t[:, :, ys + dy:ye + dy, xs + dx:xe + dx]
    return out


def head_tail_forward(feat, w, b):
    """logits (B,1,H,W) = sum_c sum_{i,j} w[0,c,i,j] * feat[c, y+i-1, x+j-1] + b   (cross-correlation, zero padding 1)."""
    out = torch.zeros(feat.shape[0], 1, feat.shape[2], feat.shape[3], dtype=feat.dtype)
    for i in range(3):
        for j in range(3):
            out += (_shift(feat, i - 1, j - 1) * w[0, :, i, j].view(1, -1, 1, 1)).sum(dim=1, keepdim=True)
    return out + b.view(1, 1, 1, 1)


def head_tail_backward(feat, w, g):
    """g = dL/dlogits (B,1,H,W) -> dfeat (B,C,H,W), dw (1,C,3,3), db (1)."""
    dfeat = torch.zeros_like(feat)
    dw = torch.zeros_like(w)
    for i in range(3):
        for j in range(3):
            dfeat += _shift(g, 1 - i, 1 - j) * w[0, :, i, j].view(1, -1, 1, 1)
            dw[0, :, i, j] = (_shift(feat, i - 1, j - 1) * g).sum(dim=(0, 2, 3))
    return dfeat, dw, g.sum().view(1)


# ------------------------------------------------------------------ ProbMaskGater (train mode) with explicit uniforms
def gate_forward(p_raw, u1, u2, *, mode="gumbel", tau=1.0, p_min=0.0, threshold=0.5):
    """probmaskgater.py:73-95.  Returns (out, saved)."""
    p = p_raw.float().clamp(0.0, 1.0)
    if p_min > 0:
        p = torch.maximum(p, torch.tensor(p_min, dtype=p.dtype))
    if mode == "deterministic":
        return p, None
    U1 = u1.clamp(1e-6, 1 - 1e-6)
    U2 = u2.clamp(1e-6, 1 - 1e-6)
    gn = -torch.log(-torch.log(U1)) - (-torch.log(-torch.log(U2)))
    pc = p.clamp(1e-6, 1.0 - 1e-6)
    logits = torch.log(pc) - torch.log1p(-pc)
    soft = torch.sigmoid((logits + gn) / tau)
    out = soft if mode == "gumbel" else (soft > threshold).float()
    return out, (p_raw, p, soft)


def gate_backward(gout, saved, *, tau=1.0, p_min=0.0):
    """d out / d p_raw: through sigmoid, the logit (zero where its clamp is active), the p_min maximum and the [0,1] clamp;
    hard_st passes the soft gradient (straight-through)."""
    p_raw, p, soft = saved
    dlogit = gout * soft * (1 - soft) / tau
    inside = (p > 1e-6) & (p < 1 - 1e-6)
    dp = torch.where(inside, dlogit / (p * (1 - p)), torch.zeros_like(p))
    if p_min > 0:
        dp = torch.where(p_raw.float().clamp(0, 1) >= p_min, dp, torch.zeros_like(dp))
    return torch.where((p_raw >= 0) & (p_raw <= 1), dp, torch.zeros_like(dp))


# ------------------------------------------------------------------ Philox4x32-10
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32(counter, key, rounds=10):
    """counter: 4 uint32, key: 2 uint32 -> 4 uint32 (numpy scalars or arrays, broadcast)."""
    c = [np.asarray(v, dtype=np.uint64) for v in counter]
    k = [np.asarray(v, dtype=np.uint64) for v in key]
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = np.uint64(_M0) * c[0]
        p1 = np.uint64(_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [(hi1 ^ c[1] ^ k[0]) & mask, lo1, (hi0 ^ c[3] ^ k[1]) & mask, lo0]
        k = [(k[0] + np.uint64(_W0)) & mask, (k[1] + np.uint64(_W1)) & mask]
    return [v.astype(np.uint32) for v in c]


def gate_uniforms(n, seed, offset):
    """The uniforms the CUDA gate kernel uses for element i: Philox4x32-10, key = (seed lo, seed hi), counter = (i lo, i hi, offset lo,
    offset hi); u1 = (r0 + 0.5) * 2^-32, u2 = (r1 + 0.5) * 2^-32 computed in fp32 (so u is in (0,1), never 0 or 1... up to rounding)."""
    i = np.arange(n, dtype=np.uint64)
    r = philox4x32([i & np.uint64(0xFFFFFFFF), i >> np.uint64(32), np.uint64(offset & 0xFFFFFFFF), np.uint64(offset >> 32)],
                   [np.uint64(seed & 0xFFFFFFFF), np.uint64(seed >> 32)])
    to_u = lambda v: ((v.astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -32)).astype(np.float32)  # noqa: E731
    return to_u(r[0]), to_u(r[1])


# ------------------------------------------------------------------ collate
def collate_masks(per_sample):
    """per_sample: list over the batch of (h_i, w_i) arrays of ONE pyramid stride -> (B,1,Hmax,Wmax) zero-padded (dataset.py:149-169)."""
    H = max(a.shape[-2] for a in per_sample)
    W = max(a.shape[-1] for a in per_sample)
    out = np.zeros((len(per_sample), 1, H, W), dtype=np.float32)
    for i, a in enumerate(per_sample):
        out[i, 0, : a.shape[-2], : a.shape[-1]] = a
    return out
