"""Forward-hook manager that applies `MaskGuidedCBAM` at the P3/P4/P5 neck outputs
(layers 15/18/21 of yolov8.yaml -- ultralytics/cfg/models/v8/yolov8.yaml:39,43,47 -- and of
the reference's configs/models/yolov8_cbam.yaml:47,51,55) of an Ultralytics detection model.

The reference injects the block through the model YAML (yolov8_cbam.yaml:67-72); its only
hooks are read-only capture hooks keyed by `m.i` (mga_yolo/model/validator.py:197,230,251).
This manager offers both graph semantics:

  semantics="layer_output" (default, the north-star wording): a forward hook REPLACES the
      output of layers 15/18/21, so the refined P3 also feeds layer 16 -> P4 -> P5
      (BaseModel._predict_once stores the hooked value in y[], ultralytics/nn/tasks.py:184-190).
  semantics="detect_input": a forward PRE-hook on the Detect head rewrites its 3-tensor
      input list; the PAN path keeps the raw features.  This is exactly the YAML graph
      (yolov8_cbam.yaml:78) and mirrors the reference's own Detect pre-hook (validator.py:222-230).

Hard requirements that shaped it:
  * the CBAM blocks are registered sub-modules of `model` (`model.mga_cbam`), so they are in
    parameters(), optimizer groups, DDP buckets, state_dict and ModelEMA's deepcopy;
  * hooks are picklable module-level callables holding references only to objects inside the
    model tree, so deepcopy()/pickle keep hook -> block wiring consistent in the copy.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Union

import torch
import torch.nn as nn

from .mask_ops import MaskUtils
from .module import MaskGuidedCBAM

MaskArg = Union[None, torch.Tensor, Sequence[Optional[torch.Tensor]], Dict[str, Optional[torch.Tensor]]]


class _MaskState:
    """The per-batch masks.  ONE state object is shared by every deep copy of the slot that owns it, so masks set through
    the manager also reach ModelEMA's copy (ultralytics/utils/torch_utils.py:752) and any validation / checkpoint copy."""

    __slots__ = ("full", "per_level", "by_stride", "keep", "fired")

    def __init__(self):
        self.full: Optional[torch.Tensor] = None  # (B,H,W) uint8 image-resolution mask
        self.per_level: Dict[str, Optional[torch.Tensor]] = {}
        self.by_stride: Dict[int, torch.Tensor] = {}
        self.keep = False   # True: masks survive the forward that used them (set_masks(..., persistent=True))
        self.fired = 0      # hooked levels served since the last set_masks

    def clear(self) -> None:
        self.full, self.per_level, self.by_stride, self.fired = None, {}, {}, 0


class _MaskSlot(nn.Module):
    """Per-batch mask holder living inside the model tree (no parameters, nothing in state_dict)."""

    def __init__(self, levels: Sequence[str], resize: str, state: Optional[_MaskState] = None):
        super().__init__()
        self.levels = tuple(levels)
        self.resize = resize
        self.state = state if state is not None else _MaskState()

    # kept for callers that poke the slot directly
    @property
    def full(self):
        return self.state.full

    @property
    def per_level(self):
        return self.state.per_level

    def clear(self) -> None:
        self.state.clear()

    def _downsampled(self, stride: int) -> torch.Tensor:
        st = self.state
        if stride in st.by_stride:
            return st.by_stride[stride]
        code = {"nearest": 0, "area": 1, "maxpool": 2}[self.resize]
        full = st.full
        from . import ops
        if stride in (8, 16, 32) and ops.masks_multi_supported(*full.shape[-2:]):
            # one kernel reads every mask once and emits the three pyramid strides (dataset.py:95-103 as a batch op)
            outs = torch.ops.mga.masks_multi(full, code, 0.0, False, True)
            for s_, o in zip((8, 16, 32), outs):
                st.by_stride[s_] = o.unsqueeze(1)
        else:
            st.by_stride[stride] = torch.ops.mga.mask_downsample(full, stride, code, 0.0, False, True).unsqueeze(1)
        return st.by_stride[stride]

    def mask_for(self, level: str, feat: torch.Tensor) -> Optional[torch.Tensor]:
        st = self.state
        try:
            if level in st.per_level:
                m = st.per_level[level]
                if m is not None and (m.shape[0] != feat.shape[0] or tuple(m.shape[-2:]) != tuple(feat.shape[-2:])):
                    raise RuntimeError(f"stale or mismatched mask for level {level}: mask {tuple(m.shape)} vs feature {tuple(feat.shape)}; "
                                       "call set_masks() for every batch")
                return m
            if st.full is None:
                return None
            H, W = feat.shape[-2:]
            fH, fW = st.full.shape[-2:]
            if st.full.shape[0] != feat.shape[0]:
                raise RuntimeError(f"stale masks: batch {st.full.shape[0]} was set, the model runs batch {feat.shape[0]}; call set_masks() per batch")
            stride = -(-fH // H)  # the reference's sizes are ceil(H / s) (mask_utils.py:84-85)
            if stride < 1 or -(-fH // stride) != H or -(-fW // stride) != W:
                raise RuntimeError(f"mask of size {(fH, fW)} does not reduce to feature size {(H, W)} by an integer stride")
            return self._downsampled(stride)
        finally:
            st.fired += 1
            if st.fired >= len(self.levels) and not st.keep:
                st.clear()  # consumed: a forward without a fresh set_masks() runs mask-free instead of re-using old masks

    def __deepcopy__(self, memo):  # copies share the live state (masks are transient and never part of a checkpoint)
        return _MaskSlot(self.levels, self.resize, self.state)

    def __getstate__(self):
        st = dict(self.__dict__)
        st["state"] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        if self.__dict__.get("state") is None:
            self.__dict__["state"] = _MaskState()


class _OutputHook:
    """forward hook: layer output -> refined output"""

    def __init__(self, block: MaskGuidedCBAM, slot: _MaskSlot, level: str):
        self.block, self.slot, self.level = block, slot, level

    def __call__(self, module, args, output):
        if not isinstance(output, torch.Tensor):
            return output
        mask = self.slot.mask_for(self.level, output)
        return self.block(output if mask is None else [output, mask])


class _DetectPreHook:
    """forward pre-hook on the detection head: refine each entry of its input list"""

    def __init__(self, blocks: nn.ModuleDict, slot: _MaskSlot, levels: Sequence[str]):
        self.blocks, self.slot, self.levels = blocks, slot, tuple(levels)

    def __call__(self, module, args):
        feats = args[0]
        if not isinstance(feats, (list, tuple)) or len(feats) != len(self.levels):
            return None
        out = []
        for lvl, f in zip(self.levels, feats):
            mask = self.slot.mask_for(lvl, f)
            out.append(self.blocks[lvl](f if mask is None else [f, mask]))
        return (out,) + tuple(args[1:])


def _layers_of(model: nn.Module) -> nn.Module:
    inner = model
    for _ in range(3):  # YOLO wrapper -> DetectionModel -> nn.Sequential
        if isinstance(inner, nn.Sequential):
            return inner
        nxt = getattr(inner, "model", None)
        if nxt is None:
            break
        inner = nxt
    if isinstance(inner, nn.Sequential):
        return inner
    raise TypeError("model does not expose an nn.Sequential of layers under .model")


def _out_channels(layer: nn.Module) -> Optional[int]:
    for path in (("cv2", "conv"), ("cv3", "conv"), ("conv",), ()):
        m = layer
        try:
            for a in path:
                m = getattr(m, a)
        except AttributeError:
            continue
        if isinstance(m, nn.Conv2d):
            return m.out_channels
    convs = [m for m in layer.modules() if isinstance(m, nn.Conv2d)]
    return convs[-1].out_channels if len(convs) == 1 else None


class MGAHookManager:
    """Attach / detach mask-guided CBAM at given layer indices.

        mgr = MGAHookManager(det_model, target_layers=("15", "18", "21"), reduction_ratio=16,
                             sam_cam_fusion="multiply", mga_pyramid_fusion="add")
        mgr.register()
        mgr.set_masks(binary_masks_uint8)   # per batch: (B,H,W)/(B,1,H,W) image-size mask, or per-level list
        preds = det_model(images)
        mgr.remove()
    """

    ATTR = "mga_cbam"

    def __init__(
        self,
        model: nn.Module,
        target_layers: Iterable[Union[str, int]] = ("15", "18", "21"),
        reduction_ratio: int = 16,
        sam_cam_fusion: str = "multiply",
        mga_pyramid_fusion: str = "add",
        *,
        channels: Optional[Sequence[int]] = None,
        semantics: str = "layer_output",
        mask_resize: str = "nearest",
        use_sigmoid_mask: bool = False,
        spatial_k: int = 7,
    ) -> None:
        if semantics not in ("layer_output", "detect_input"):
            raise ValueError("semantics must be 'layer_output' or 'detect_input'")
        if mask_resize not in ("nearest", "area", "maxpool"):
            raise ValueError("mask_resize must be 'nearest', 'area' or 'maxpool'")
        self.levels: List[str] = [str(t) for t in target_layers]
        self.semantics = semantics
        self.layers = _layers_of(model)
        self.owner = model if not hasattr(model, "model") or isinstance(model, nn.Sequential) else model
        while not isinstance(self.owner, nn.Module):  # e.g. ultralytics.YOLO wrapper
            self.owner = self.owner.model
        n = len(self.layers)
        for lvl in self.levels:
            if not (0 <= int(lvl) < n):
                raise IndexError(f"target layer {lvl} outside model with {n} layers")
        if channels is None:
            channels = [self._infer_channels(int(lvl)) for lvl in self.levels]
        if len(channels) != len(self.levels):
            raise ValueError("channels must have one entry per target layer")
        existing = getattr(self.owner, self.ATTR, None)
        if isinstance(existing, nn.ModuleDict) and set(existing.keys()) >= set(self.levels):
            self.blocks = existing  # re-attach to a model that already owns its blocks (e.g. loaded checkpoint)
        else:
            ref = next((p for p in self.owner.parameters()), None)
            self.blocks = nn.ModuleDict({
                lvl: MaskGuidedCBAM(c, reduction_ratio=reduction_ratio, spatial_k=spatial_k, use_sigmoid_mask=use_sigmoid_mask,
                                    sam_cam_fusion=sam_cam_fusion, mga_pyramid_fusion=mga_pyramid_fusion)
                for lvl, c in zip(self.levels, channels)
            })
            if ref is not None:
                self.blocks.to(ref.device)
            setattr(self.owner, self.ATTR, self.blocks)
        self.slot = _MaskSlot(self.levels, mask_resize)
        setattr(self.owner, self.ATTR + "_masks", self.slot)
        self._handles: list = []

    def _infer_channels(self, idx: int) -> int:
        head = self.layers[-1]
        f = getattr(head, "f", None)
        if isinstance(f, (list, tuple)) and idx in f and hasattr(head, "cv2"):
            first = head.cv2[list(f).index(idx)][0]
            conv = getattr(first, "conv", first)
            if isinstance(conv, nn.Conv2d):
                return conv.in_channels
        c = _out_channels(self.layers[idx])
        if c is None:
            raise ValueError(f"cannot infer the channel count of layer {idx}; pass channels=[...]")
        return c

    # -- registration -----------------------------------------------------------------
    def register(self) -> "MGAHookManager":
        if self._handles:
            return self
        if self.semantics == "layer_output":
            for lvl in self.levels:
                h = self.layers[int(lvl)].register_forward_hook(_OutputHook(self.blocks[lvl], self.slot, lvl))
                self._handles.append(h)
        else:
            head = self.layers[-1]
            self._handles.append(head.register_forward_pre_hook(_DetectPreHook(self.blocks, self.slot, self.levels)))
        return self

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    @property
    def registered(self) -> bool:
        return bool(self._handles)

    # -- per-batch masks ---------------------------------------------------------------
    def set_masks(self, masks: MaskArg, *, model: Optional[nn.Module] = None, persistent: bool = False) -> None:
        """Masks of the NEXT forward.  They are consumed by that forward (every hooked level served once) unless
        `persistent=True`.  `model=` addresses the slot of another model object (e.g. one rebuilt from a checkpoint);
        deep copies of the managed model (ModelEMA) share the manager's own slot state and need nothing."""
        slot = self.slot if model is None else self.slot_of(model)
        st = slot.state
        st.clear()
        st.keep = bool(persistent)
        if masks is None:
            return
        if isinstance(masks, torch.Tensor):
            m = masks
            if m.dim() == 4:
                m = m[:, 0]
            st.full = (m > 0).to(torch.uint8).contiguous()
            return
        if isinstance(masks, dict):
            st.per_level = {str(k): v for k, v in masks.items()}
            return
        masks = list(masks)
        if len(masks) != len(self.levels):
            raise ValueError(f"expected {len(self.levels)} per-level masks, got {len(masks)}")
        st.per_level = dict(zip(self.levels, masks))

    @classmethod
    def slot_of(cls, model: nn.Module) -> _MaskSlot:
        inner = model
        for _ in range(4):
            slot = getattr(inner, cls.ATTR + "_masks", None)
            if isinstance(slot, _MaskSlot):
                return slot
            inner = getattr(inner, "model", None)
            if inner is None:
                break
        raise AttributeError("model carries no MGA mask slot (was MGAHookManager constructed on it?)")

    def alphas(self) -> Dict[str, float]:
        return {lvl: float(self.blocks[lvl].alpha.detach()) for lvl in self.levels}

    def __enter__(self):
        return self.register()

    def __exit__(self, *exc):
        self.remove()
