"""CPU restatement of the reference's MaskECA block (forward + closed-form backward).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing under mga_yolo_b200/).  Follows
/root/reference/mga_yolo/nn/modules/masked_eca.py:
  eca_kernel_size        :43-52   k = odd(clamp(int(|floor(log2 C) / gamma + b|), k_min, k_max))
  MaskECA._pool          :139-164 masked average (sigmoid mask, clamp_min(sum m, eps), GAP fall-back when mean(m) < thr)
  MaskECA.forward        :166-193 y -> conv1d(1,1,k, pad k//2, no bias) over the channel axis -> sigmoid -> g = 1 + softplus(beta)(w - 0.5) -> x * g
Pinned against fixtures produced by running that class (oracle/gen_golden_next.py -> tests/golden/eca_*.npz) in
tests/test_oracle_golden.py.  The backward is the closed form the CUDA kernels implement, not autograd.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F


def eca_kernel_size(channels: int, gamma: float = 2.0, b: float = 1.0, k_min: int = 3, k_max: int = 15) -> int:
    """masked_eca.py:43-52."""
    if channels <= 0:
        return k_min
    k = int(abs((channels.bit_length() - 1) / gamma + b))
    k = max(k_min, min(k, k_max))
    return k if k % 2 == 1 else k + 1


@dataclass
class EcaSaved:
    x: torch.Tensor
    m: Optional[torch.Tensor]      # (B,S) mask after sigmoid, None without a mask
    y: torch.Tensor                # (B,C) pooled descriptor
    A: torch.Tensor                # (B,C) masked average before the blend
    use: torch.Tensor              # (B)
    den: torch.Tensor              # (B)
    msum: torch.Tensor             # (B)
    w: torch.Tensor                # (B,C) sigmoid(conv1d(y))
    g: torch.Tensor                # (B,C) gate
    w1d: torch.Tensor              # (k)
    beta: torch.Tensor
    use_sigmoid_mask: bool
    eps: float
    mask_shape: Optional[tuple]


def eca_forward(x, mask, w1d, beta, *, use_sigmoid_mask=True, tiny_thr=1e-4, eps=1e-6, feature_dtype=None):
    """x (B,C,H,W), mask (B,1,H,W) | (B,H,W) | None, w1d (k) = conv1d.weight.view(-1), beta ().  Computes in x.dtype (fp64 in the tests);
    feature_dtype rounds x first (the 16-bit parity cases: fp32 oracle on rounded inputs)."""
    if feature_dtype is not None and feature_dtype != x.dtype:
        x = x.to(feature_dtype).to(x.dtype)
    B, C, H, W = x.shape
    S = H * W
    xf = x.reshape(B, C, S)
    G = xf.mean(dim=2)
    if mask is None:
        m = None
        y, A = G, G
        use = torch.zeros(B, dtype=x.dtype)
        den = torch.ones(B, dtype=x.dtype)
        msum = torch.zeros(B, dtype=x.dtype)
    else:
        mshape = tuple(mask.shape)
        m = mask.reshape(B, S).to(x.dtype)
        if use_sigmoid_mask:
            m = torch.sigmoid(m)
        msum = m.sum(dim=1)
        use = (msum / S >= tiny_thr).to(x.dtype)
        den = msum.clamp_min(eps)
        A = torch.einsum("bcs,bs->bc", xf, m) / den[:, None]
        y = A * use[:, None] + G * (1 - use[:, None])
    k = w1d.numel()
    z = F.conv1d(y[:, None, :], w1d.view(1, 1, k).to(x.dtype), padding=k // 2)[:, 0]
    w = torch.sigmoid(z)
    alpha = F.softplus(beta.to(x.dtype))
    g = 1 + alpha * (w - 0.5)
    out = x * g[:, :, None, None]
    sv = EcaSaved(x, m, y, A, use, den, msum, w, g, w1d.to(x.dtype), beta.to(x.dtype), use_sigmoid_mask, eps,
                  None if mask is None else tuple(mask.shape))
    return out, sv


def eca_backward(gout, sv: EcaSaved):
    """Closed-form gradients: dx, dmask (None without a mask), conv1d.weight (1,1,k), beta."""
    x = sv.x
    B, C, H, W = x.shape
    S = H * W
    gf = gout.to(x.dtype).reshape(B, C, S)
    xf = x.reshape(B, C, S)
    dg = (gf * xf).sum(dim=2)                                  # (B,C)
    alpha = F.softplus(sv.beta)
    dalpha = (dg * (sv.w - 0.5)).sum()
    dbeta = torch.sigmoid(sv.beta) * dalpha
    dz = alpha * dg * sv.w * (1 - sv.w)                        # (B,C)
    k = sv.w1d.numel()
    pad = k // 2
    ypad = F.pad(sv.y, (pad, pad))
    dw = torch.stack([(dz * ypad[:, j:j + C]).sum() for j in range(k)])
    dy = F.conv1d(dz[:, None, :], sv.w1d.flip(0).view(1, 1, k), padding=pad)[:, 0]   # transpose of the forward conv
    if sv.m is None:
        dx = gf * sv.g[:, :, None] + (dy / S)[:, :, None]
        dmask = None
    else:
        use, den = sv.use[:, None], sv.den[:, None]
        cA = use * dy / den                                     # (B,C)
        cG = (1 - use) * dy / S
        dx = gf * sv.g[:, :, None] + cA[:, :, None] * sv.m[:, None, :] + cG[:, :, None]
        passthru = (sv.msum >= sv.eps).to(x.dtype)[:, None]     # clamp_min backward
        dm = torch.einsum("bc,bcs->bs", cA, xf) - (cA * sv.A * passthru).sum(dim=1, keepdim=True)
        if sv.use_sigmoid_mask:
            dm = dm * sv.m * (1 - sv.m)
        dmask = dm.reshape(sv.mask_shape)
    return {"dx": dx.reshape(B, C, H, W), "dmask": dmask, "conv1d.weight": dw.view(1, 1, k), "beta": dbeta}
