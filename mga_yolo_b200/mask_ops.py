"""GPU mirror of the reference's mask downsample helpers (mga_yolo/utils/mask_utils.py).

Same names, argument meaning and environment switches as `MaskUtils.downsample_mask`
(:64-141) and `MaskUtils.downsample_mask_prob` (:14-48), but the input is a CUDA uint8
tensor -- one mask (H,W) or a batch (B,H,W) -- and the work is one kernel launch instead
of cv2 calls in dataloader workers.  Bit-exact with cv2 (tests/test_gpu_mask.py).
"""
from __future__ import annotations

import os
from typing import Sequence

import torch

from . import _lib, ops  # noqa: F401  (ops registers torch.ops.mga.*)


def _binary_u8(mask: torch.Tensor) -> torch.Tensor:
    if not mask.is_cuda:
        raise RuntimeError("mga_yolo_b200.mask_ops works on CUDA tensors only (no CPU fallback; the CPU path is cv2 in the reference)")
    return mask if mask.dtype == torch.uint8 else (mask > 0).to(torch.uint8)


class MaskUtils:
    @staticmethod
    def downsample_mask(mask: torch.Tensor, stride: int) -> torch.Tensor:
        """uint8 {0,1} (…,H,W) -> (…,ceil(H/s),ceil(W/s)); method from env MGA_MASK_METHOD
        ('nearest' | 'area' | 'maxpool' | default connectivity-preserving = block max + 3x3 close),
        MGA_MASK_BRIDGE=0 disables the close, MGA_MASK_THRESH is the 'area' threshold."""
        method = os.getenv("MGA_MASK_METHOD", "skeleton_bresenham").lower()
        bridge = os.getenv("MGA_MASK_BRIDGE", "1") not in {"0", "false", "False"}
        thresh = float(os.getenv("MGA_MASK_THRESH", "0.0"))
        mask = _binary_u8(mask)
        if stride <= 1:
            return mask
        if method == "nearest":
            return torch.ops.mga.mask_downsample(mask, stride, _lib.DS_NEAREST, 0.0, False, False)
        if method == "area":
            return torch.ops.mga.mask_downsample(mask, stride, _lib.DS_AREA, thresh, bridge, False)
        if method == "maxpool":
            return torch.ops.mga.mask_downsample(mask, stride, _lib.DS_MAXPOOL, 0.0, False, False)
        if method == "pyrdown" or os.getenv("MGA_SKELETON_STRICT", "0").lower() in {"1", "true", "yes"}:
            raise NotImplementedError(f"mask method {method!r} (pyrDown / strict skeleton) is outside the B200 hot path")
        # default: non-strict skeleton_bresenham == occupancy (block max) + optional 3x3 close
        return torch.ops.mga.mask_downsample(mask, stride, _lib.DS_MAXPOOL, 0.0, bridge, False)

    @staticmethod
    def downsample_mask_prob(mask: torch.Tensor, stride: int, method: str = "area") -> torch.Tensor:
        """float32 in [0,1]; 'avgpool' = exact block mean, 'nearest', 'area' (uint8 INTER_AREA -> {0,1})."""
        if stride <= 1:
            return mask.float()
        mask = _binary_u8(mask)
        code = {"avgpool": _lib.DS_AVGPOOL, "nearest": _lib.DS_NEAREST}.get(method, _lib.DS_AREA_RAW)
        return torch.ops.mga.mask_downsample(mask, stride, code, 0.0, False, True)

    @staticmethod
    def _method_args(prob: bool):
        """(method code, thresh, close3x3, out_float) of downsample_mask / downsample_mask_prob for the current environment."""
        if prob:
            method = os.getenv("MGA_MASK_METHOD", "area")
            code = {"avgpool": _lib.DS_AVGPOOL, "nearest": _lib.DS_NEAREST}.get(method, _lib.DS_AREA_RAW)
            return code, 0.0, False, True
        method = os.getenv("MGA_MASK_METHOD", "skeleton_bresenham").lower()
        bridge = os.getenv("MGA_MASK_BRIDGE", "1") not in {"0", "false", "False"}
        thresh = float(os.getenv("MGA_MASK_THRESH", "0.0"))
        if method == "nearest":
            return _lib.DS_NEAREST, 0.0, False, False
        if method == "area":
            return _lib.DS_AREA, thresh, bridge, False
        if method == "maxpool":
            return _lib.DS_MAXPOOL, 0.0, False, False
        if method == "pyrdown" or os.getenv("MGA_SKELETON_STRICT", "0").lower() in {"1", "true", "yes"}:
            return None
        return _lib.DS_MAXPOOL, 0.0, bridge, False

    @staticmethod
    def collate_masks(per_sample: Sequence[torch.Tensor]) -> torch.Tensor:
        """Zero-pad to the largest (h, w) of the batch and stack: (B,1,H,W) float32, one pyramid stride per call -- the
        `masks_multi` branch of MGADataset.collate_fn (mga_yolo/data/dataset.py:149-169) on device tensors (uint8 or float32)."""
        return torch.ops.mga.collate_masks(list(per_sample))

    @staticmethod
    def masks_multi(bin_masks: torch.Tensor, strides: Sequence[int] = (8, 16, 32), prob: bool = False) -> list:
        """Batch form of MGADataset.__getitem__'s loop (mga_yolo/data/dataset.py:95-103):
        (B,H,W) uint8 -> [ (B,1,Hs,Ws) for s in strides ].  Letterboxed sizes (H, W multiples of 32) with the default
        strides take ONE kernel that reads every mask once (torch.ops.mga.masks_multi); anything else goes stride by stride."""
        args = MaskUtils._method_args(prob)
        if tuple(strides) == (8, 16, 32) and bin_masks.dim() == 3 and args is not None and ops.masks_multi_supported(*bin_masks.shape[-2:]):
            code, thresh, close, out_f = args
            outs = torch.ops.mga.masks_multi(_binary_u8(bin_masks), code, thresh, close, out_f)
            return [o.unsqueeze(-3) for o in outs]
        method = os.getenv("MGA_MASK_METHOD", "area")
        outs = []
        for s in strides:
            ds = MaskUtils.downsample_mask_prob(bin_masks, s, method) if prob else MaskUtils.downsample_mask(bin_masks, s)
            outs.append(ds.unsqueeze(-3))
        return outs
