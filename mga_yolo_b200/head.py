"""`MGAMaskHead` -- the mask-logit producer next to the block.

Mirror of the reference head (mga_yolo/nn/modules/segmentation.py:57-118): Conv1x1(in -> hidden, no bias) -> BatchNorm | channel-last
LayerNorm -> activation -> (Dropout2d) -> Conv3x3(hidden -> out).  Same constructor, same state_dict (`proj.*`, `head.weight`,
`head.bias`), same Kaiming / ones / zeros initialisation order.  The 1x1 projection is a dense GEMM and stays a library convolution
(cuDNN; out of the hot-path scope).  The TAIL -- the 3x3 convolution that produces the logits the CBAM block
reads and that receives dL/dmask back from it -- runs in the CUDA library when out_channels == 1 and the tensor is on a GPU
(`torch.ops.mga.head_tail_fwd/bwd`), in fp32 output precision whatever the feature dtype.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import next_ops


@dataclass
class MGAMaskHeadConfig:
    in_channels: int
    hidden_channels: int
    out_channels: int = 1
    norm: Optional[str] = "bn"
    act: type = nn.SiLU
    dropout: float = 0.0


class ChannelLastLayerNorm(nn.Module):
    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.ln = nn.LayerNorm(num_channels, eps=eps)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.ln(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class MGAMaskHead(nn.Module):
    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int = 1, norm: Optional[str] = "bn", act: type = nn.SiLU,
                 dropout: float = 0.0) -> None:
        super().__init__()
        self.cfg = MGAMaskHeadConfig(in_channels, hidden_channels, out_channels, norm, act, dropout)
        layers = [nn.Conv2d(in_channels, hidden_channels, kernel_size=1, bias=False)]
        if norm == "bn":
            layers.append(nn.BatchNorm2d(hidden_channels))
        elif norm == "ln":
            layers.append(ChannelLastLayerNorm(hidden_channels))
        if act is not None:
            layers.append(act())
        if dropout and dropout > 0:
            layers.append(nn.Dropout2d(p=dropout))
        self.proj = nn.Sequential(*layers)
        self.head = nn.Conv2d(hidden_channels, out_channels, kernel_size=3, padding=1, bias=True)  # parameter container of the tail
        for m in self.modules():  # segmentation.py:97-105
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.BatchNorm2d, nn.LayerNorm)):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        feat = self.proj(x)
        if self.cfg.out_channels == 1 and feat.is_cuda and feat.dtype in (torch.float32, torch.bfloat16, torch.float16) and feat.shape[1] <= 256:
            return next_ops.head_tail(feat, self.head.weight, self.head.bias)
        return self.head(feat)  # multi-channel masks / CPU shape probes: the library convolution, like the reference

    def extra_repr(self) -> str:
        c = self.cfg
        return f"in={c.in_channels}, hidden={c.hidden_channels}, out={c.out_channels}, norm={c.norm}, act={c.act.__name__ if c.act else None}, dropout={c.dropout}"
