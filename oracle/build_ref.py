"""Recipe for oracle/_ref: the reference's OWN implementation of the hot path, importable without the rest of the repo.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by mga_yolo_b200/).  The reference is pure Python: this recipe lays its
hot-path source files, unmodified and byte for byte, from where they lie under /root/reference into oracle/_ref/ (git-ignored,
but shipped to the GPU box with the gpurun snapshot) inside a skeleton package whose only hand-written parts are empty
`__init__.py` files and a two-line stand-in for `ultralytics.utils.LOGGER` (a logging.Logger; mask_utils.py:7,
segmentation.py:30 import it).  No reference source is ever committed to this repository.

    python oracle/build_ref.py            # in the authoring container (needs /root/reference)

Used by: tests/ (oracle pinning, reference-vs-CUDA parity at full size on the GPU box), bench.py --impl reference
(`cpu_baseline.kind = "reference"`) and bench.py's `gpu_eager_baseline`.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DST = HERE / "_ref"
SRC = Path("/root/reference")

# reference file -> the hot-path role it plays
FILES = {
    "mga_yolo/nn/modules/masked_cbam.py": "MaskCBAM (8a rows a1-a7)",
    "mga_yolo/nn/modules/probmaskgater.py": "ProbMaskGater (a8)",
    "mga_yolo/utils/mask_utils.py": "MaskUtils.downsample_mask[_prob] (a9, a10)",
    "mga_yolo/utils/mask_downsample.py": "default connectivity-preserving method (a9)",
    "mga_yolo/nn/modules/segmentation.py": "MGAMaskHead (8f-1)",
    "mga_yolo/nn/modules/masked_eca.py": "MaskECA (8f-4)",
    "mga_yolo/nn/modules/masked_spade.py": "MaskSPADE (8f-4)",
}
STUBS = {
    "mga_yolo/__init__.py": "",
    "mga_yolo/nn/__init__.py": "",
    "mga_yolo/nn/modules/__init__.py": "",
    "mga_yolo/utils/__init__.py": "",
    "mga_yolo/external/__init__.py": "",
    "mga_yolo/external/ultralytics/__init__.py": "",
    "mga_yolo/external/ultralytics/ultralytics/__init__.py": "",
    "mga_yolo/external/ultralytics/ultralytics/utils/__init__.py": "import logging\n\nLOGGER = logging.getLogger('ultralytics')\n",
}


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref is in place (built now or already there), False when it cannot be built here."""
    if not SRC.exists():
        return (DST / "MANIFEST.json").exists()
    if DST.exists():
        shutil.rmtree(DST)
    manifest = {}
    for rel, text in STUBS.items():
        p = DST / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(text)
    for rel, role in FILES.items():
        src = SRC / rel
        if not src.exists():
            raise FileNotFoundError(src)
        dst = DST / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = {"role": role, "sha256": hashlib.sha256(dst.read_bytes()).hexdigest(), "bytes": dst.stat().st_size}
    (DST / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files laid out under {DST}")
    return True


def available() -> bool:
    return (DST / "MANIFEST.json").exists()


def load():
    """Import the reference classes from oracle/_ref.  Returns a namespace with MaskCBAM, ProbMaskGater, MaskUtils, MGAMaskHead,
    MaskECA, MaskSPADE.  Raises RuntimeError when oracle/_ref is absent or the full reference tree is already imported."""
    import importlib
    import types

    if not available():
        raise RuntimeError("oracle/_ref is not built: run `python oracle/build_ref.py` in the authoring container")
    have = sys.modules.get("mga_yolo")
    if have is not None and str(DST) not in str(getattr(have, "__file__", "") or getattr(have, "__path__", "")):
        # the whole reference checkout is imported in this process (tests in the authoring container): use it as is
        root = None
    else:
        root = str(DST)
        if root not in sys.path:
            sys.path.insert(0, root)
    ns = types.SimpleNamespace()
    ns.MaskCBAM = importlib.import_module("mga_yolo.nn.modules.masked_cbam").MaskCBAM
    ns.ProbMaskGater = importlib.import_module("mga_yolo.nn.modules.probmaskgater").ProbMaskGater
    ns.MaskUtils = importlib.import_module("mga_yolo.utils.mask_utils").MaskUtils
    seg = importlib.import_module("mga_yolo.nn.modules.segmentation")
    ns.MGAMaskHead = seg.MGAMaskHead
    ns.MaskECA = importlib.import_module("mga_yolo.nn.modules.masked_eca").MaskECA
    ns.MaskSPADE = importlib.import_module("mga_yolo.nn.modules.masked_spade").MaskSPADE
    ns.root = root
    return ns


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok else 1)
