"""mga_yolo_b200 -- B200-native (sm_100a) mask-guided CBAM hot path of MGA-YOLO.

Public surface (mirrors the reference's operator interface for this path):
    MaskGuidedCBAM / MaskCBAM   nn.Module drop-in for mga_yolo/nn/modules/masked_cbam.py
    MGAHookManager              forward-hook injection at layers 15/18/21
    MaskUtils                   GPU mirror of mga_yolo/utils/mask_utils.py (downsample_mask[_prob])
    install / uninstall         swap the class inside the reference's Ultralytics graph builder
    ops.mask_guided_cbam        functional form;  torch.ops.mga.{cbam_fwd,cbam_bwd,mask_downsample}
"""
from . import ops  # noqa: F401  registers torch.ops.mga.*
from .dist import FlatGradReducer, shard_range
from .hooks import MGAHookManager
from .install import install, uninstall
from .mask_ops import MaskUtils
from .module import MaskCBAM, MaskGate, MaskGuidedCBAM

__all__ = [
    "MaskGuidedCBAM", "MaskCBAM", "MaskGate", "MGAHookManager", "MaskUtils", "install", "uninstall",
    "FlatGradReducer", "shard_range", "ops",
]
__version__ = "0.1.0"
