"""Quick timing of the MaskECA kernels at one pyramid level: python tools/eca_prof.py [C H W B dtype]  (torch profiler: per-kernel device time)"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mga_yolo_b200 import MaskECA  # noqa: E402

C, H, W, B = (int(v) for v in (sys.argv[1:5] or (64, 80, 80, 64)))
dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[sys.argv[5] if len(sys.argv) > 5 else "f32"]
dev = torch.device("cuda:0")
mod = MaskECA(C).to(dev)
x = torch.randn(B, C, H, W, device=dev, dtype=dt).requires_grad_(True)
m = torch.randn(B, 1, H, W, device=dev).requires_grad_(True)
g = torch.randn(B, C, H, W, device=dev, dtype=dt)


def step():
    out = mod([x, m])
    out.backward(g)
    x.grad = None
    m.grad = None
    mod.zero_grad()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows) / 5
n = x.numel() * x.element_size()
print(f"== ECA {B}x{C}x{H}x{W} {dt}: device time {tot:.1f} us per fwd+bwd; 7N = {7 * n / 1e6:.0f} MB -> {7 * n / tot / 1e3:.0f} GB/s")
for r in rows[:10]:
    print(f"   {r.device_time_total / 5:8.1f} us  x{r.count // 5:<3d} {r.key[:100]}")
