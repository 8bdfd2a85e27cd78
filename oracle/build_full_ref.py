"""Recipe for baseline/_ref: the UNMODIFIED reference package (mga_yolo/ with its vendored Ultralytics tree) plus its model YAMLs,
laid out from /root/reference so that the full-model GPU test (tests/test_gpu_full_model.py) can build `MGAModel` on the GPU box, where
/root/reference does not exist.  `pip install --target baseline/_ref /root/reference` was tried first: it builds a wheel of the Python
modules only (no YAML package data, so `ultralytics.cfg` cannot even be imported), hence this plain copy.  baseline/_ref is git-ignored
(no reference source is committed) but travels with the gpurun snapshot.  TEST INFRASTRUCTURE ONLY.

    python oracle/build_full_ref.py
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DST = ROOT / "baseline" / "_ref"
SRC = Path("/root/reference")
KEEP = {".py", ".yaml", ".yml", ".json", ".txt", ".cfg", ".toml"}
SKIP_DIRS = {"docs", "examples", "docker", "tests", ".github", "__pycache__", "assets"}


def build(verbose: bool = True) -> bool:
    if not SRC.exists():
        return (DST / "mga_yolo" / "__init__.py").exists()
    if DST.exists():
        shutil.rmtree(DST)
    n = 0
    for top in ("mga_yolo", "configs"):
        for p in (SRC / top).rglob("*"):
            if not p.is_file() or p.suffix not in KEEP or any(part in SKIP_DIRS for part in p.relative_to(SRC).parts):
                continue
            q = DST / p.relative_to(SRC)
            q.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(p, q)
            n += 1
    if verbose:
        print(f"baseline/_ref: {n} reference files laid out under {DST}")
    return True


def available() -> bool:
    return (DST / "mga_yolo" / "__init__.py").exists() and (DST / "configs" / "models" / "yolov8_cbam.yaml").exists()


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
