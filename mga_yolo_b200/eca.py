"""`MaskECA` -- the nn.Module face of the B200 mask-guided ECA block.

Drop-in for the reference class (mga_yolo/nn/modules/masked_eca.py:69-200): same constructor, same `cfg` dataclass, same state_dict
(`conv1d.weight (1,1,k)`, `beta ()`), same `forward(x)` polymorphism (`Tensor` or `[feature, mask]`), `.alpha`, `.scale_name`.  The
math (masked average pool -> conv1d over channels -> sigmoid -> residual gate -> rescale, and its closed-form backward) is one call
into the CUDA library; the pooling front end is the CBAM block's.  CPU tensors raise: there is no compute fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, next_ops
from .module import in_shape_probe


def eca_kernel_size(channels: int, gamma: float = 2.0, b: float = 1.0, k_min: int = 3, k_max: int = 15) -> int:
    """masked_eca.py:43-52."""
    if channels <= 0:
        return k_min
    k = int(abs((channels.bit_length() - 1) / gamma + b))
    k = max(k_min, min(k, k_max))
    return k if k % 2 == 1 else k + 1


@dataclass
class MaskECAConfig:
    channels: int
    gamma: float = 2.0
    b: float = 1.0
    k_min: int = 3
    k_max: int = 15
    use_sigmoid_mask: bool = True
    tiny_mask_threshold: float = 1e-4
    eps: float = 1e-6


class MaskECA(nn.Module):
    def __init__(self, channels: int, gamma: float = 2.0, b: float = 1.0, k_min: int = 3, k_max: int = 15, use_sigmoid_mask: bool = True,
                 tiny_mask_threshold: float = 1e-4, eps: float = 1e-6) -> None:
        super().__init__()
        self.cfg = MaskECAConfig(channels, gamma, b, k_min, k_max, use_sigmoid_mask, tiny_mask_threshold, eps)
        k = eca_kernel_size(channels, gamma=gamma, b=b, k_min=k_min, k_max=k_max)
        self.conv1d = nn.Conv1d(1, 1, kernel_size=k, padding=k // 2, bias=False)  # parameter container (same init as the reference)
        self.beta = nn.Parameter(torch.tensor(0.0, dtype=torch.float32))
        self.scale_name: str = {256: "P3", 512: "P4", 1024: "P5"}.get(channels, f"C{channels}")

    @property
    def alpha(self) -> torch.Tensor:
        return F.softplus(self.beta)

    def _maybe_rebuild_conv(self, channels: int) -> None:
        """masked_eca.py:124-137: a fresh conv1d when the runtime channel count differs from the configured one."""
        if channels == self.cfg.channels:
            return
        k = eca_kernel_size(channels, gamma=self.cfg.gamma, b=self.cfg.b, k_min=self.cfg.k_min, k_max=self.cfg.k_max)
        weight = self.conv1d.weight
        self.conv1d = nn.Conv1d(1, 1, kernel_size=k, padding=k // 2, bias=False).to(device=weight.device, dtype=weight.dtype)
        self.cfg.channels = channels

    def forward(self, x: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
        if isinstance(x, (list, tuple)):
            assert len(x) == 2, "MaskECA expects [feature, mask] as inputs"
            feat, mask = x
        else:
            feat, mask = x, None
        assert isinstance(feat, torch.Tensor) and feat.dim() == 4, "feature must be (B,C,H,W)"
        if feat.device.type == "cpu":
            if in_shape_probe():
                return torch.zeros_like(feat)
            raise RuntimeError("mga_yolo_b200.MaskECA runs on CUDA tensors only (no CPU fallback); move the model to a GPU")
        self._maybe_rebuild_conv(feat.shape[1])
        if mask is not None:
            if not mask.is_floating_point():
                mask = mask.to(torch.float32)
            elif mask.dtype == torch.float64:
                mask = mask.float()
        flags = _lib.SIGMOID_MASK if self.cfg.use_sigmoid_mask else 0
        return next_ops.mask_eca(feat, mask, self.conv1d.weight, self.beta, flags=flags, tiny_mask_thr=self.cfg.tiny_mask_threshold,
                                 eps=self.cfg.eps)

    def extra_repr(self) -> str:
        c = self.cfg
        return (f"C={c.channels}, gamma={c.gamma}, b={c.b}, k_min={c.k_min}, k_max={c.k_max}, sigmoid_mask={c.use_sigmoid_mask}, "
                f"tiny_thr={c.tiny_mask_threshold}, alpha={float(self.alpha.detach())}, scale='{self.scale_name}'")
