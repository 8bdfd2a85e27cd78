"""`MaskSPADE` -- the nn.Module face of the B200 mask-conditioned SPADE block.

Drop-in for the reference class (mga_yolo/nn/modules/masked_spade.py:41-151): same constructor, same `cfg` dataclass, same state_dict
(`shared.0.{weight,bias}`, `conv_gamma.{weight,bias}`, `conv_beta.{weight,bias}`; the affine-free InstanceNorm2d holds nothing), same
initialisation order, same `forward(x)` polymorphism (`Tensor` or `[feature, mask]`), `.scale_name`.

Split of the work: the FEATURE side -- instance statistics, normalisation, `gamma * xhat + beta` and the closed-form backward of all
of it (4 N / 5 N elements of HBM traffic, the part that scales with B*C*H*W) -- is one call into the CUDA library
(`mga_spade_forward` / `mga_spade_backward`, csrc/spade_ops.cu).  The MASK branch that produces gamma / beta (Conv3x3 1->hidden, ReLU,
two Conv3x3 hidden->C, masked_spade.py:78-84) is dense convolution work and runs on the library convolutions, as in the reference.
`norm_type="bn"` (cross-sample statistics) is not mirrored.  CPU tensors raise: there is no compute fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import next_ops
from .module import in_shape_probe


@dataclass
class MaskSPADEConfig:
    channels: int
    hidden: int = 64
    mask_channels: int = 1
    norm_type: str = "in"
    use_sigmoid_mask: bool = True
    eps: float = 1e-6


class MaskSPADE(nn.Module):
    def __init__(self, channels: int, hidden: int = 64, mask_channels: int = 1, norm_type: str = "in", use_sigmoid_mask: bool = True,
                 eps: float = 1e-6) -> None:
        super().__init__()
        if norm_type.lower() == "bn":
            raise NotImplementedError("mga_yolo_b200.MaskSPADE mirrors norm_type='in' (the reference default, masked_spade.py:57,72-75); "
                                      "batch statistics are not built")
        self.cfg = MaskSPADEConfig(channels, hidden, mask_channels, norm_type, use_sigmoid_mask, eps)
        self.norm = nn.InstanceNorm2d(channels, affine=False, eps=eps)  # attribute kept for parity of the module tree; holds no state
        in_mc = max(1, mask_channels)
        self.shared = nn.Sequential(nn.Conv2d(in_mc, hidden, kernel_size=3, padding=1, bias=True), nn.ReLU(inplace=True))
        self.conv_gamma = nn.Conv2d(hidden, channels, kernel_size=3, padding=1, bias=True)
        self.conv_beta = nn.Conv2d(hidden, channels, kernel_size=3, padding=1, bias=True)
        self.scale_name: str = {256: "P3", 512: "P4", 1024: "P5"}.get(channels, f"C{channels}")
        for m in self.modules():  # masked_spade.py:92-100 (same visiting order => same weights under the same seed)
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    @staticmethod
    def _prep_mask(mask: torch.Tensor, target_hw, use_sigmoid: bool) -> torch.Tensor:
        """masked_spade.py:102-112."""
        if mask.dim() == 3:
            mask = mask.unsqueeze(1)
        if tuple(mask.shape[-2:]) != tuple(target_hw):
            mask = F.interpolate(mask, size=tuple(target_hw), mode="bilinear", align_corners=False)
        return mask.sigmoid() if use_sigmoid else mask

    def forward(self, x: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
        if isinstance(x, (list, tuple)):
            assert len(x) == 2, "MaskSPADE expects [feature, mask] as inputs"
            feat, mask = x
        else:
            feat, mask = x, None
        assert isinstance(feat, torch.Tensor) and feat.dim() == 4, "feature must be (B,C,H,W)"
        if feat.device.type == "cpu":
            if in_shape_probe():
                return torch.zeros_like(feat)
            raise RuntimeError("mga_yolo_b200.MaskSPADE runs on CUDA tensors only (no CPU fallback); move the model to a GPU")
        if mask is None:
            return next_ops.spade_modulate(feat, None, None, self.cfg.eps)
        mask = self._prep_mask(mask, feat.shape[-2:], self.cfg.use_sigmoid_mask)
        h = self.shared(mask.to(self.conv_gamma.weight.dtype))
        gamma, beta = self.conv_gamma(h), self.conv_beta(h)
        if gamma.dtype not in (feat.dtype, torch.float32):  # masked_spade.py:139-141
            gamma, beta = gamma.to(feat.dtype), beta.to(feat.dtype)
        elif beta.dtype != gamma.dtype:
            beta = beta.to(gamma.dtype)
        return next_ops.spade_modulate(feat, gamma, beta, self.cfg.eps)

    def extra_repr(self) -> str:
        c = self.cfg
        return (f"C={c.channels}, hidden={c.hidden}, maskC={c.mask_channels}, norm={c.norm_type}, "
                f"sigmoid_mask={c.use_sigmoid_mask}, scale='{self.scale_name}'")
