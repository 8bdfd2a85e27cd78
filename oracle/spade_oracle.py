"""CPU restatement of the reference's MaskSPADE block (forward + closed-form backward of the feature side).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing under mga_yolo_b200/).  Follows
/root/reference/mga_yolo/nn/modules/masked_spade.py:
  norm                   :72-75    nn.InstanceNorm2d(C, affine=False, eps): per (sample, channel) mean and BIASED variance over H*W
  _prep_mask             :102-112  (B,H,W) -> (B,1,H,W); bilinear resize (align_corners=False) when the size differs; sigmoid
  forward                :114-144  xhat = norm(x); no mask -> xhat; h = ReLU(conv3x3(mask)); gamma, beta = conv3x3(h); y = gamma * xhat + beta
Pinned against fixtures produced by running that class (oracle/gen_golden_spade.py -> tests/golden/spade_*.npz) in
tests/test_oracle_golden.py.  `modulate_backward` is the closed form the CUDA kernels implement, not autograd; the mask branch
(dense convolutions, library work on both sides) is restated with F.conv2d and differentiated by autograd.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def instance_stats(x: torch.Tensor, eps: float):
    """mean and rstd per (b, c) over H*W (biased variance, eps inside the root) -- what InstanceNorm2d(affine=False) uses."""
    B, C, H, W = x.shape
    xf = x.reshape(B, C, H * W)
    mean = xf.mean(dim=2)
    var = ((xf - mean[:, :, None]) ** 2).mean(dim=2)
    return mean, 1.0 / torch.sqrt(var + eps)


def modulate_forward(x, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float = 1e-6):
    """y = gamma * xhat + beta (xhat when gamma is None); returns (y, saved)."""
    mean, rstd = instance_stats(x, eps)
    xhat = (x - mean[:, :, None, None]) * rstd[:, :, None, None]
    y = xhat if gamma is None else gamma * xhat + beta
    return y, {"xhat": xhat, "rstd": rstd, "gamma": gamma}


def modulate_backward(g, saved) -> Dict[str, Optional[torch.Tensor]]:
    """d gamma = g xhat; d beta = g; dx = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat)), dxhat = g gamma."""
    xhat, rstd, gamma = saved["xhat"], saved["rstd"], saved["gamma"]
    dxhat = g if gamma is None else g * gamma
    m1 = dxhat.mean(dim=(2, 3), keepdim=True)
    m2 = (dxhat * xhat).mean(dim=(2, 3), keepdim=True)
    dx = rstd[:, :, None, None] * (dxhat - m1 - xhat * m2)
    return {"dx": dx, "dgamma": None if gamma is None else g * xhat, "dbeta": None if gamma is None else g}


def prep_mask(mask, target_hw, use_sigmoid: bool):
    if mask.dim() == 3:
        mask = mask.unsqueeze(1)
    if tuple(mask.shape[-2:]) != tuple(target_hw):
        mask = F.interpolate(mask, size=tuple(target_hw), mode="bilinear", align_corners=False)
    return mask.sigmoid() if use_sigmoid else mask


def mask_branch(mask, params: Dict[str, torch.Tensor], target_hw, use_sigmoid: bool = True):
    """gamma, beta (B,C,H,W) from the mask: masked_spade.py:132-137."""
    m = prep_mask(mask, target_hw, use_sigmoid)
    h = F.relu(F.conv2d(m, params["shared.0.weight"], params["shared.0.bias"], padding=1))
    gamma = F.conv2d(h, params["conv_gamma.weight"], params["conv_gamma.bias"], padding=1)
    beta = F.conv2d(h, params["conv_beta.weight"], params["conv_beta.bias"], padding=1)
    return gamma, beta


PARAM_KEYS = ("shared.0.weight", "shared.0.bias", "conv_gamma.weight", "conv_gamma.bias", "conv_beta.weight", "conv_beta.bias")


def spade_forward_backward(x, mask, params, g, *, use_sigmoid_mask=True, eps=1e-6):
    """Whole block: out, dx, dmask and the six parameter gradients.  Feature side by the closed form above, mask branch by autograd."""
    if mask is None:
        out, sv = modulate_forward(x, None, None, eps)
        return {"out": out, "dx": modulate_backward(g, sv)["dx"]}
    mask = mask.clone().requires_grad_(True)
    ps = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    gamma, beta = mask_branch(mask, ps, x.shape[-2:], use_sigmoid_mask)
    out, sv = modulate_forward(x, gamma.detach(), beta.detach(), eps)
    gr = modulate_backward(g, sv)
    torch.autograd.backward([gamma, beta], [gr["dgamma"], gr["dbeta"]])
    res = {"out": out, "dx": gr["dx"], "dmask": mask.grad, "gamma": gamma.detach(), "beta": beta.detach()}
    for k in PARAM_KEYS:
        res[k] = ps[k].grad
    return res
