"""cl_bwd device time with and without the mask gradient (hook-manager / ground-truth masks do not need it): python tools/nodmask_prof.py [C H W B dtype]"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mga_yolo_b200 import MaskGuidedCBAM  # noqa: E402

C, H, W, B = (int(v) for v in (sys.argv[1:5] or (64, 80, 80, 64)))
dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[sys.argv[5] if len(sys.argv) > 5 else "f32"]
dev = torch.device("cuda:0")
mod = MaskGuidedCBAM(C).to(dev)
g = torch.randn(B, C, H, W, device=dev, dtype=dt)
for need in (True, False):
    x = torch.randn(B, C, H, W, device=dev, dtype=dt).requires_grad_(True)
    m = torch.randn(B, 1, H, W, device=dev).requires_grad_(need)

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
    for r in prof.key_averages():
        if "cl_bwd" in r.key or "cl_fwd" in r.key:
            print(f"mask grad {need}: {r.device_time_total / 5:8.1f} us  {r.key[:60]}")
