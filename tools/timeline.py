"""Debug: per-phase %globaltimer stamps of the fused kernels (thread 0 of every CTA)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import LevelPlan, WORKLOADS, DT  # noqa: E402
from mga_yolo_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.mga_debug_timeline.argtypes = [C.c_void_p]
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
which = sys.argv[2] if len(sys.argv) > 2 else "fwd"
levels, B, dtname, _ = WORKLOADS[wl]
import os
B = int(os.environ.get("MGA_TL_BATCH", B))
dev = torch.device("cuda:0")
for li, (Cc, H, W) in enumerate(levels):
    flat = torch.zeros(LevelPlan.n_params(Cc), device=dev)
    pl = LevelPlan(None, Cc, H, W, B, DT[dtname], 0, dev, li, flat, 0)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        pl.fwd(st); pl.bwd(st)
    torch.cuda.synchronize()
    buf = torch.zeros(16 * 4096, dtype=torch.int64, device=dev)
    lib.mga_debug_timeline(buf.data_ptr())
    if which == "fwd":
        pl.fwd(st)
    else:
        pl.fwd(st); torch.cuda.synchronize(); buf.zero_(); pl.bwd(st)
    torch.cuda.synchronize()
    lib.mga_debug_timeline(None)
    t = buf.cpu().numpy().reshape(-1, 16)
    t = t[t[:, 0] > 0]
    nst = int((t[0] > 0).sum())
    t0 = t[:, 0].min()
    rel = (t[:, :nst] - t0) / 1e3
    d = np.diff(t[:, :nst], axis=1) / 1e3
    print(f"level P{3+li} C={Cc} {H}x{W}: {len(t)} CTAs, kernel span {rel.max():.1f} us")
    print("  phase durations (us) median:", np.round(np.median(d, axis=0), 2))
    print("  phase durations (us) max   :", np.round(d.max(axis=0), 2))
    late = t[:, 0] > np.percentile(t[:, 0], 60)  # CTAs of the later waves (steady state)
    if late.sum() > 8:
        print("  phase durations (us) median, later waves:", np.round(np.median(d[late], axis=0), 2), " lifetime", round(float(np.median(rel[late, nst - 1] - rel[late, 0])), 2))
    print("  CTA start (us) pctl 0/50/100:", np.round(np.percentile(rel[:, 0], [0, 50, 100]), 1),
          " CTA lifetime median:", round(float(np.median(rel[:, nst - 1] - rel[:, 0])), 2))
