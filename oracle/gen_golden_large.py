"""Generate tests/golden/cbam_large_{p3,p4,p5}.npz by RUNNING THE REFERENCE MaskCBAM (authoring container only) on the three
YOLOv8n neck shapes at batch 2.  Inputs are regenerated from seeds by tests/_golden.large_inputs (not stored); stored are the
parameters, 4096 sampled elements of out / dx (fp32 and fp64 reference), the whole dmask and every parameter gradient.

    PYTHONPATH=/root/reference:. PYTHONDONTWRITEBYTECODE=1 YOLO_CONFIG_DIR=/tmp/ulcfg python oracle/gen_golden_large.py
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
sys.dont_write_bytecode = True
for p in ("/root/reference", str(Path(__file__).resolve().parent.parent)):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

from tests._golden import GOLDEN, LARGE_CASES, large_inputs


def main():
    for key in ("MGA_PROB_MODE", "MGA_PROB_APPROACH"):
        os.environ.pop(key, None)
    from mga_yolo.nn.modules.masked_cbam import MaskCBAM

    for tag in LARGE_CASES:
        x, mask, g, beta, seed, idx = large_inputs(tag)
        C = x.shape[1]
        torch.manual_seed(seed)
        mod = MaskCBAM(C)
        with torch.no_grad():
            mod.beta.fill_(beta)
        rec = {"p." + n: t.detach().numpy().copy() for n, t in mod.state_dict().items()}
        for suffix, dt in (("", torch.float32), ("_f64", torch.float64)):
            m = mod.double() if dt == torch.float64 else mod.float()
            xi = x.to(dt).clone().requires_grad_(True)
            mi = mask.to(dt).clone().requires_grad_(True)
            for prm in m.parameters():
                prm.grad = None
            out = m([xi, mi])
            out.backward(g.to(dt))
            rec["out" + suffix] = out.detach().reshape(-1)[idx].numpy().copy()
            rec["dx" + suffix] = xi.grad.reshape(-1)[idx].numpy().copy()
            rec["dmask" + suffix] = mi.grad.numpy().copy()
            rec["out_absmax" + suffix] = np.array(out.detach().abs().max().item())
            rec["dx_absmax" + suffix] = np.array(xi.grad.abs().max().item())
            for n, prm in m.named_parameters():
                rec["d." + n + suffix] = prm.grad.detach().numpy().copy()
        np.savez_compressed(GOLDEN / f"cbam_large_{tag}.npz", **rec)
        print(tag, {k: v.shape for k, v in rec.items() if k.startswith(("out", "dx", "dmask"))})


if __name__ == "__main__":
    main()
