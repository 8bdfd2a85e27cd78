"""CPU oracle for the mask-guided CBAM hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path (mga_yolo_b200/) never does; it
fails loudly when the CUDA library is missing.

Parity status
-------------
* sam_cam_fusion="multiply" + mga_pyramid_fusion="add" (the only behaviour the
  reference implements): PINNED.  tests/golden/cbam_*.npz were produced by running
  the reference class itself (oracle/gen_golden.py imports
  mga_yolo/nn/modules/masked_cbam.py from /root/reference) and
  tests/test_oracle_golden.py checks this restatement against them, forward and
  all eight gradients, in fp32 and fp64.
* every other fusion mode: "parity unpinned" -- no reference source or test
  defines them.  They are build-side definitions made
  from the same primitives (channel scale s, spatial map a, mask m, alpha).

What is restated (reference file:line)
--------------------------------------
  masked average pool   mga_yolo/nn/modules/masked_cbam.py:87-102
  masked max pool       mga_yolo/nn/modules/masked_cbam.py:104-121
  channel attention     mga_yolo/nn/modules/masked_cbam.py:123-130
  spatial attention     mga_yolo/nn/modules/masked_cbam.py:132-148
  alpha residual        mga_yolo/nn/modules/masked_cbam.py:150-152,166-171
  eval-mode mask gate   mga_yolo/nn/modules/probmaskgater.py:73-83
The backward is NOT autograd: it is the closed form of SURVEY.md section 8a so the
CUDA kernels can be checked term by term.

All functions are dtype generic (float32 / float64 torch CPU tensors).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SAM_CAM_MODES = ("multiply", "add")
PYRAMID_MODES = ("add", "multiply")


@dataclass
class CbamParams:
    """Tensors named after the reference state_dict (masked_cbam.py:53-64)."""

    w1: torch.Tensor  # cam_mlp.0.weight (h, C)
    b1: torch.Tensor  # cam_mlp.0.bias   (h,)
    w2: torch.Tensor  # cam_mlp.2.weight (C, h)
    b2: torch.Tensor  # cam_mlp.2.bias   (C,)
    wsam: torch.Tensor  # sam_conv.weight (1, 3, k, k)
    beta: torch.Tensor  # ()

    def to(self, dtype=None, device=None) -> "CbamParams":
        return CbamParams(*(t.detach().to(dtype=dtype, device=device) for t in (self.w1, self.b1, self.w2, self.b2, self.wsam, self.beta)))

    @staticmethod
    def from_state_dict(sd: Dict[str, torch.Tensor]) -> "CbamParams":
        return CbamParams(
            sd["cam_mlp.0.weight"], sd["cam_mlp.0.bias"], sd["cam_mlp.2.weight"], sd["cam_mlp.2.bias"],
            sd["sam_conv.weight"], sd["beta"],
        )


@dataclass
class CbamSaved:
    """Everything the closed-form backward needs."""

    tensors: Dict[str, torch.Tensor] = field(default_factory=dict)
    flags: Dict[str, object] = field(default_factory=dict)


def _very_low(ref_dtype: torch.dtype) -> float:
    # masked_cbam.py:115 -- finfo(x.dtype).min of the FEATURE dtype
    return float(torch.finfo(ref_dtype).min)


def cbam_forward(
    x: torch.Tensor,
    mask: Optional[torch.Tensor],
    p: CbamParams,
    *,
    use_sigmoid_mask: bool = True,
    tiny_mask_thr: float = 1e-4,
    eps: float = 1e-6,
    sam_cam_fusion: str = "multiply",
    mga_pyramid_fusion: str = "add",
    gate_clamp: bool = False,
    feature_dtype: Optional[torch.dtype] = None,
):
    """Forward of the block.  Returns (out, CbamSaved).

    `feature_dtype` is the dtype the real feature map has (bf16 / fp16 / fp32); it
    only selects the "very low" sentinel of the masked max (masked_cbam.py:115).
    `gate_clamp` applies the eval-mode ProbMaskGater (probmaskgater.py:77,82-83)
    in front of the block: the raw mask is clamped to [0,1] first.
    """
    assert sam_cam_fusion in SAM_CAM_MODES and mga_pyramid_fusion in PYRAMID_MODES
    assert x.dim() == 4
    B, C, H, W = x.shape
    S = H * W
    dt = x.dtype
    xf = x.reshape(B, C, S)
    sv = CbamSaved()
    has_mask = mask is not None

    if has_mask:
        raw = mask.reshape(B, S).to(dt)  # (B,H,W) and (B,1,H,W) are both accepted (masked_cbam.py:81-85)
        pre_gate = raw
        if gate_clamp:
            raw = raw.clamp(0.0, 1.0)
        m = torch.sigmoid(raw) if use_sigmoid_mask else raw
        msum = m.sum(dim=1)  # (B,)
        use = (msum / S >= tiny_mask_thr).to(dt)  # masked_cbam.py:97-98
        den = msum.clamp_min(eps)  # masked_cbam.py:99
        A = (xf * m[:, None, :]).sum(dim=2) / den[:, None]  # masked_cbam.py:100
        G = xf.mean(dim=2)  # masked_cbam.py:101
        avg = A * use[:, None] + G * (1.0 - use[:, None])
        valid = m > 0.5  # masked_cbam.py:116
        low = _very_low(feature_dtype or dt)
        xm = torch.where(valid[:, None, :], xf, torch.full((), low, dtype=dt, device=x.device))
        mraw, amax_hw = xm.max(dim=2)  # first max index (adaptive_max_pool2d scan order)
        dead = torch.isclose(mraw, torch.full((), low, dtype=dt, device=x.device))  # masked_cbam.py:120
        mx = torch.where(dead, G, mraw)
    else:
        m = torch.zeros(B, S, dtype=dt, device=x.device)
        pre_gate = None
        G = xf.mean(dim=2)
        avg = G
        mx, amax_hw = xf.max(dim=2)
        dead = torch.zeros(B, C, dtype=torch.bool, device=x.device)
        use = torch.zeros(B, dtype=dt, device=x.device)
        den = torch.ones(B, dtype=dt, device=x.device)
        msum = torch.zeros(B, dtype=dt, device=x.device)
        A = G

    # shared MLP on both descriptors, summed (masked_cbam.py:128): b2 enters twice
    ha = torch.relu(avg @ p.w1.t() + p.b1)
    hm = torch.relu(mx @ p.w1.t() + p.b1)
    z = ha @ p.w2.t() + p.b2 + hm @ p.w2.t() + p.b2
    s = torch.sigmoid(z)  # (B,C)

    q = s if sam_cam_fusion == "multiply" else torch.ones_like(s)
    y1 = xf * q[:, :, None]
    pmax, amax_c = y1.max(dim=1)  # (B,S); ties -> lowest channel index
    pavg = y1.mean(dim=1)
    cat = torch.stack([pmax, pavg, m], dim=1).reshape(B, 3, H, W)  # plane order masked_cbam.py:146
    k = p.wsam.shape[-1]
    pre = F.conv2d(cat, p.wsam, padding=k // 2)
    a = torch.sigmoid(pre).reshape(B, S)

    if sam_cam_fusion == "multiply":
        gate = s[:, :, None] * a[:, None, :]
    else:
        gate = s[:, :, None] + a[:, None, :]
    alpha = F.softplus(p.beta)
    k0 = (1.0 - alpha) if mga_pyramid_fusion == "add" else torch.zeros((), dtype=dt, device=x.device)
    k1 = alpha
    out = (xf * (k0 + k1 * gate)).reshape(B, C, H, W)

    sv.tensors = dict(
        x=xf, m=m, use=use, den=den, msum=msum, A=A, avg=avg, mx=mx, dead=dead, amax_hw=amax_hw,
        ha=ha, hm=hm, s=s, q=q, amax_c=amax_c, cat=cat, a=a, gate=gate, alpha=alpha, k0=k0, k1=k1,
        pre_gate=pre_gate,
    )
    sv.flags = dict(
        has_mask=has_mask, use_sigmoid_mask=use_sigmoid_mask, eps=eps, sam_cam_fusion=sam_cam_fusion,
        mga_pyramid_fusion=mga_pyramid_fusion, gate_clamp=gate_clamp, shape=(B, C, H, W), mask_shape=None if mask is None else tuple(mask.shape),
    )
    return out, sv


def cbam_backward(g: torch.Tensor, p: CbamParams, sv: CbamSaved) -> Dict[str, torch.Tensor]:
    """Closed-form gradients of `cbam_forward` (generalised to the
    build-side fusion modes).  Returns dx, dmask (None without a mask) and the six
    parameter gradients keyed like the reference state_dict."""
    t, fl = sv.tensors, sv.flags
    B, C, H, W = fl["shape"]
    S = H * W
    xf, s, a, q, m = t["x"], t["s"], t["a"], t["q"], t["m"]
    dt = xf.dtype
    multiply = fl["sam_cam_fusion"] == "multiply"
    gf = g.reshape(B, C, S).to(dt)
    gx = gf * xf

    # alpha = softplus(beta); d softplus = sigmoid
    if fl["mga_pyramid_fusion"] == "add":
        dalpha = (gx * (t["gate"] - 1.0)).sum()
    else:
        dalpha = (gx * t["gate"]).sum()
    dbeta = torch.sigmoid(p.beta) * dalpha

    dgate = t["k1"] * gx
    if multiply:
        da = (dgate * s[:, :, None]).sum(dim=1)
        ds = (dgate * a[:, None, :]).sum(dim=2)
    else:
        da = dgate.sum(dim=1)
        ds = dgate.sum(dim=2)
    dpre = (da * a * (1.0 - a)).reshape(B, 1, H, W)
    k = p.wsam.shape[-1]
    dwsam = torch.nn.grad.conv2d_weight(t["cat"], p.wsam.shape, dpre, padding=k // 2)
    dcat = torch.nn.grad.conv2d_input(t["cat"].shape, p.wsam, dpre, padding=k // 2).reshape(B, 3, S)

    onehot_c = F.one_hot(t["amax_c"], C).permute(0, 2, 1).to(dt)  # (B,C,S)
    dy1 = dcat[:, 1:2, :] / C + onehot_c * dcat[:, 0:1, :]
    dx = gf * (t["k0"] + t["k1"] * t["gate"]) + q[:, :, None] * dy1
    if multiply:
        ds = ds + (xf * dy1).sum(dim=2)

    dz = ds * s * (1.0 - s)
    ha, hm = t["ha"], t["hm"]
    dw2 = dz.t() @ (ha + hm)
    db2 = 2.0 * dz.sum(dim=0)
    dha = (dz @ p.w2) * (ha > 0).to(dt)
    dhm = (dz @ p.w2) * (hm > 0).to(dt)
    dw1 = dha.t() @ t["avg"] + dhm.t() @ t["mx"]
    db1 = (dha + dhm).sum(dim=0)
    davg = dha @ p.w1
    dmx = dhm @ p.w1

    onehot_hw = F.one_hot(t["amax_hw"], S).to(dt)  # (B,C,S)
    dmask = None
    if fl["has_mask"]:
        use, den = t["use"], t["den"]
        dead = t["dead"].to(dt)
        dG = (1.0 - use)[:, None] * davg + dead * dmx
        dA = use[:, None] * davg
        dx = dx + dG[:, :, None] / S + dA[:, :, None] * m[:, None, :] / den[:, None, None]
        dx = dx + ((1.0 - dead) * dmx)[:, :, None] * onehot_hw
        clamp_pass = (t["msum"] >= fl["eps"]).to(dt)  # clamp_min backward
        dm = (dA[:, :, None] * (xf - (t["A"] * clamp_pass[:, None])[:, :, None])).sum(dim=1) / den[:, None]
        dm = dm + dcat[:, 2, :]
        if fl["use_sigmoid_mask"]:
            dm = dm * m * (1.0 - m)
        if fl["gate_clamp"]:
            pg = t["pre_gate"]
            dm = dm * ((pg >= 0.0) & (pg <= 1.0)).to(dt)
        dmask = dm.reshape(fl["mask_shape"])
    else:
        dx = dx + davg[:, :, None] / S + dmx[:, :, None] * onehot_hw

    return {
        "dx": dx.reshape(B, C, H, W),
        "dmask": dmask,
        "cam_mlp.0.weight": dw1,
        "cam_mlp.0.bias": db1,
        "cam_mlp.2.weight": dw2,
        "cam_mlp.2.bias": db2,
        "sam_conv.weight": dwsam,
        "beta": dbeta,
    }


def cbam_forward_general(x, mask, p: CbamParams, *, sam_cam_fusion="multiply", mga_pyramid_fusion="add", fuse_sam_cam=None,
                         fuse_pyramid=None, **kw):
    """All nine fusion combinations, built from the same primitives (build-side definitions,
    reference parity unpinned except multiply/add).  Differentiable with torch autograd (used in fp64 as the checker).
    fuse_* = (weight (C,2C,1,1), bias (C)) of the extra 1x1 convolution of a `concat` mode."""
    B, C, H, W = x.shape
    inner = "multiply" if sam_cam_fusion == "multiply" else "add"  # where the spatial gate is computed from: x*s or x
    _, sv = cbam_forward(x, mask, p, sam_cam_fusion=inner, mga_pyramid_fusion="add", **kw)
    s = sv.tensors["s"][:, :, None, None]
    a = sv.tensors["a"].reshape(B, 1, H, W)
    if sam_cam_fusion == "multiply":
        R = x * s * a
    elif sam_cam_fusion == "add":
        R = x * s + x * a
    else:
        R = F.conv2d(torch.cat([x * s, x * a], dim=1), fuse_sam_cam[0], fuse_sam_cam[1])
    alpha = F.softplus(p.beta)
    if mga_pyramid_fusion == "add":
        return x + alpha * (R - x)
    if mga_pyramid_fusion == "multiply":
        return alpha * R
    return F.conv2d(torch.cat([x, R], dim=1), fuse_pyramid[0], fuse_pyramid[1])


def cbam_forward_autograd(x, mask, p: CbamParams, **kw):
    """Same forward, but built so torch autograd can differentiate it -- used as the CPU
    baseline in bench.py (it costs what the reference's eager path costs: one ATen call
    per line) and as a cross-check of the closed-form backward."""
    out, _ = cbam_forward(x, mask, p, **kw)
    return out


def default_params(C: int, r: int = 16, k: int = 7, *, seed: int = 0, beta: float = 0.0, dtype=torch.float32) -> CbamParams:
    """Deterministic parameters with the reference's shapes (masked_cbam.py:53-64).
    Same init family as nn.Linear / nn.Conv2d defaults (uniform +-1/sqrt(fan_in))."""
    gen = torch.Generator().manual_seed(seed)
    h = max(1, C // r)

    def u(shape, fan_in):
        bound = 1.0 / fan_in ** 0.5
        return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    return CbamParams(
        w1=u((h, C), C), b1=u((h,), C), w2=u((C, h), h), b2=u((C,), h),
        wsam=u((1, 3, k, k), 3 * k * k), beta=torch.tensor(beta, dtype=dtype),
    )
