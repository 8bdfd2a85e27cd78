"""ncu target / quick timing of the MaskSPADE feature-side kernels at one pyramid level: python tools/spade_prof.py [C H W B dtype]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import mga_yolo_b200  # noqa: F401  (registers torch.ops.mga.*)

C, H, W, B = (int(v) for v in (sys.argv[1:5] or (64, 80, 80, 64)))
dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[sys.argv[5] if len(sys.argv) > 5 else "f32"]
dev = torch.device("cuda:0")
x, gm, bt, g = (torch.randn(B, C, H, W, device=dev, dtype=dt) for _ in range(4))
for _ in range(3):
    _, st = torch.ops.mga.spade_fwd(x, gm, bt, 1e-6)
    torch.ops.mga.spade_bwd(g, x, gm, st, True)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for _ in range(10):
    _, st = torch.ops.mga.spade_fwd(x, gm, bt, 1e-6)
ev[1].record()
for _ in range(10):
    torch.ops.mga.spade_bwd(g, x, gm, st, True)
ev[2].record()
torch.cuda.synchronize()
n = x.numel() * x.element_size()
tf, tb = ev[0].elapsed_time(ev[1]) / 10, ev[1].elapsed_time(ev[2]) / 10
print(f"spade {B}x{C}x{H}x{W} {dt}: fwd {tf * 1e3:.1f} us = {4 * n / tf / 1e6:.0f} GB/s, bwd {tb * 1e3:.1f} us = {5 * n / tb / 1e6:.0f} GB/s")
