"""Swap the reference's `MaskCBAM` for the B200 implementation, in place.

parse_model resolves YAML class names through `globals()` of ultralytics/nn/tasks.py and
then tests `m is MaskCBAM` (tasks.py:1676-1682,1733-1739), so the patched object has to be
the very attribute of that module; MGATrainer imports the class late from
mga_yolo.nn.modules.masked_cbam (mga_yolo/model/trainer.py:286).  Call `install()` BEFORE
building the model (`YOLO(...)`/`MGAModel(...)`); `uninstall()` restores the originals.
"""
from __future__ import annotations

import sys
from typing import Dict, Tuple

from .module import MaskCBAM

_TARGETS = (
    "mga_yolo.nn.modules.masked_cbam",
    "ultralytics.nn.tasks",
    "ultralytics.nn",
    "mga_yolo.external.ultralytics.ultralytics.nn.tasks",
    "mga_yolo.external.ultralytics.ultralytics.nn",
)
_saved: Dict[Tuple[str, str], object] = {}


def install(strict: bool = False) -> list:
    """Patch every already-imported module that exposes `MaskCBAM`; returns the patched module names."""
    done = []
    for name in _TARGETS:
        mod = sys.modules.get(name)
        if mod is None or not hasattr(mod, "MaskCBAM"):
            continue
        key = (name, "MaskCBAM")
        if key not in _saved:
            _saved[key] = getattr(mod, "MaskCBAM")
        setattr(mod, "MaskCBAM", MaskCBAM)
        done.append(name)
    if strict and not any(n.endswith("nn.tasks") for n in done):
        raise RuntimeError("ultralytics.nn.tasks is not imported yet: import the reference's ultralytics first, then install()")
    return done


def uninstall() -> None:
    for (name, attr), obj in list(_saved.items()):
        mod = sys.modules.get(name)
        if mod is not None:
            setattr(mod, attr, obj)
        del _saved[(name, attr)]
