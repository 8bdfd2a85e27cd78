"""Kernel-level time breakdown (torch profiler, CUDA activities) of one fwd+bwd of the concat / multiply module at the cfg4 shapes.
usage: python tools/concat_bwd_prof.py [batch]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mga_yolo_b200 import MaskGuidedCBAM  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
for (Cc, H, W) in [(256, 80, 80), (512, 40, 40), (512, 20, 20)]:
    torch.manual_seed(Cc)
    mod = MaskGuidedCBAM(Cc, sam_cam_fusion="concat", mga_pyramid_fusion="multiply").to(dev)
    x = torch.randn(B, Cc, H, W, device=dev).bfloat16().requires_grad_(True)
    m = torch.randn(B, 1, H, W, device=dev).requires_grad_(True)
    g = torch.randn(B, Cc, H, W, device=dev).bfloat16()

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
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
    tot = sum(r.device_time_total for r in rows) / 3
    print(f"== C={Cc} {H}x{W} B={B}: total device time {tot:.0f} us per step")
    for r in rows[:16]:
        print(f"   {r.device_time_total / 3:8.1f} us  x{r.count // 3:<3d} {r.key[:110]}")
