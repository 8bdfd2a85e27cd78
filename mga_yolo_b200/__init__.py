"""mga_yolo_b200 -- B200-native (sm_100a) mask-guided CBAM hot path of MGA-YOLO.

Public surface (mirrors the reference's operator interface for this path):
    MaskGuidedCBAM / MaskCBAM   nn.Module drop-in for mga_yolo/nn/modules/masked_cbam.py
    MGAHookManager              forward-hook injection at layers 15/18/21
    MaskUtils                   GPU mirror of mga_yolo/utils/mask_utils.py (downsample_mask[_prob])
    install / uninstall         swap the class inside the reference's Ultralytics graph builder
    ops.mask_guided_cbam        functional form;  torch.ops.mga.{cbam_fwd,cbam_bwd,mask_downsample}
    MaskECA                     nn.Module drop-in for mga_yolo/nn/modules/masked_eca.py (shares the masked-pool front end)
    MaskSPADE                   nn.Module drop-in for mga_yolo/nn/modules/masked_spade.py (instance norm + modulation + backward in CUDA)
    MGAMaskHead                 mirror of mga_yolo/nn/modules/segmentation.py whose 3x3 logit tail runs in the CUDA library
    next_ops.gate_sample        train-mode ProbMaskGater (Philox noise contract);  MaskUtils.collate_masks  zero-pad collate
"""
from . import ops  # noqa: F401  registers torch.ops.mga.*
from . import next_ops  # noqa: F401  MaskECA / head tail / gate sampling / collate ops
from .dist import FlatGradReducer, shard_range
from .hooks import MGAHookManager
from .install import install, uninstall
from .mask_ops import MaskUtils
from .module import MaskCBAM, MaskGate, MaskGuidedCBAM
from .eca import MaskECA
from .head import MGAMaskHead
from .spade import MaskSPADE

__all__ = [
    "MaskGuidedCBAM", "MaskCBAM", "MaskGate", "MGAHookManager", "MaskUtils", "install", "uninstall",
    "FlatGradReducer", "shard_range", "ops", "next_ops", "MaskECA", "MGAMaskHead", "MaskSPADE",
]
__version__ = "0.1.0"
