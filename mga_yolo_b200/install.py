"""Swap the reference's `MaskCBAM` (and its neighbours `MaskECA`, `MaskSPADE`, `MGAMaskHead`) for the B200 implementations, in place.

parse_model resolves YAML class names through `globals()` of ultralytics/nn/tasks.py and
then tests `m is MaskCBAM` (tasks.py:1676-1682,1733-1739), so the patched object has to be
the very attribute of that module; MGATrainer imports the class late from
mga_yolo.nn.modules.masked_cbam (mga_yolo/model/trainer.py:286).  Call `install()` BEFORE
building the model (`YOLO(...)`/`MGAModel(...)`); `uninstall()` restores the originals.
"""
from __future__ import annotations

import sys
from typing import Dict, Tuple

import functools

from .eca import MaskECA
from .head import MGAMaskHead
from .module import MaskCBAM, shape_probe
from .spade import MaskSPADE

# YAML class name -> replacement (ultralytics/nn/tasks.py:1724 `m is MGAMaskHead`, :1733 `m is MaskECA or ... m is MaskCBAM`)
_CLASSES = {"MaskCBAM": MaskCBAM, "MaskECA": MaskECA, "MaskSPADE": MaskSPADE, "MGAMaskHead": MGAMaskHead}
_TARGETS = (
    "mga_yolo.nn.modules.masked_cbam",
    "mga_yolo.nn.modules.masked_eca",
    "mga_yolo.nn.modules.masked_spade",
    "mga_yolo.nn.modules.segmentation",
    "mga_yolo.model.model",  # `isinstance(m, MGAMaskHead)` against its import-time global decides which layers feed "seg" (model.py:11,217-220)
    "ultralytics.nn.tasks",
    "ultralytics.nn",
    "mga_yolo.external.ultralytics.ultralytics.nn.tasks",
    "mga_yolo.external.ultralytics.ultralytics.nn",
)
_saved: Dict[Tuple[str, str], object] = {}


def _wrap_builder(cls) -> None:
    """DetectionModel.__init__ runs a CPU forward of zeros to read the output strides (ultralytics/nn/tasks.py:418-426)
    before the model can be on a GPU: run it inside shape_probe(), where the block answers with shapes only."""
    init = cls.__dict__.get("__init__")
    if init is None or getattr(init, "_mga_shape_probe", False):
        return

    @functools.wraps(init)
    def __init__(self, *args, **kwargs):
        with shape_probe():
            return init(self, *args, **kwargs)

    __init__._mga_shape_probe = True
    __init__._mga_original = init
    cls.__init__ = __init__


def _unwrap_builder(cls) -> None:
    init = cls.__dict__.get("__init__")
    if init is not None and getattr(init, "_mga_shape_probe", False):
        cls.__init__ = init._mga_original


def install(strict: bool = False, classes=("MaskCBAM", "MaskECA", "MaskSPADE", "MGAMaskHead")) -> list:
    """Patch every already-imported module that exposes one of `classes`; returns the patched module names.
    The graph builder's `DetectionModel.__init__` (the base of MGAModel, mga_yolo/model/model.py:40) is wrapped so that
    its CPU stride probe passes through the block as a shape-only call."""
    done = []
    for name in _TARGETS:
        mod = sys.modules.get(name)
        if mod is None:
            continue
        hit = False
        for attr in classes:
            if getattr(mod, attr, None) is None:  # (tasks.py sets the name to None when its import failed)
                continue
            key = (name, attr)
            if key not in _saved:
                _saved[key] = getattr(mod, attr)
            setattr(mod, attr, _CLASSES[attr])
            hit = True
        if not hit:
            continue
        if name.endswith("nn.tasks") and isinstance(getattr(mod, "DetectionModel", None), type):
            _wrap_builder(mod.DetectionModel)
        done.append(name)
    if strict and not any(n.endswith("nn.tasks") for n in done):
        raise RuntimeError("ultralytics.nn.tasks is not imported yet: import the reference's ultralytics first, then install()")
    return done


def uninstall() -> None:
    for (name, attr), obj in list(_saved.items()):
        mod = sys.modules.get(name)
        if mod is not None:
            setattr(mod, attr, obj)
            if name.endswith("nn.tasks") and isinstance(getattr(mod, "DetectionModel", None), type):
                _unwrap_builder(mod.DetectionModel)
        del _saved[(name, attr)]
