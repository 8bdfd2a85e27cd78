"""Kernel-level durations (torch profiler) of the one-pass mask pyramid.  usage: python tools/mask_prof.py [batch] [imgsz]"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mga_yolo_b200 import MaskUtils  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 640
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
bufs = [torch.from_numpy((rng.random((B, N, N)) > 0.7).astype(np.uint8)).to(dev) for _ in range(6)]
for i in range(3):
    MaskUtils.masks_multi(bufs[i])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(6):
        MaskUtils.masks_multi(bufs[i])
    torch.cuda.synchronize()
for r in sorted(prof.key_averages(), key=lambda r: -r.device_time_total):
    print(f"{r.device_time_total / 6:8.1f} us  x{r.count // 6:<2d} {r.key[:120]}")
