"""Per-phase %globaltimer stamps of the persistent kernels (thread 0 of every CTA, the CTA's SECOND item = steady state).
usage: MGA_LIBNAME=libmga_cbam_tuning.so python tools/pf_timeline.py [cfg2|cfg3] [fwd|bwd] [batch]"""
import ctypes as C
import os
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
if len(sys.argv) > 3:
    B = int(sys.argv[3])
NAMES = {"fwd": ["prologue", "wait x", "pool", "barrier1", "combine", "MLP", "max/mean", "barrier2", "halo", "conv", "rescale+store"],
         "bwd": ["context", "reduce1(x,g)", "barrier1", "halo+convT+dW", "reduce2", "barrier2", "MLP bwd", "dx+store"]}[which]
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
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if which == "fwd":
        a.record(); pl.fwd(st); b.record()
    else:
        pl.fwd(st); torch.cuda.synchronize(); buf.zero_(); a.record(); pl.bwd(st); b.record()
    torch.cuda.synchronize()
    lib.mga_debug_timeline(None)
    t = buf.cpu().numpy().reshape(-1, 16)[:, :len(NAMES) + 1]
    t = t[(t > 0).all(axis=1)]
    if len(t) == 0:
        print(f"P{3+li}: no CTA had a second item"); continue
    d = np.diff(t, axis=1) / 1e3
    print(f"P{3+li} C={Cc} {H}x{W} B={B}: call {a.elapsed_time(b)*1e3:.1f} us; {len(t)} CTAs stamped; item time median {np.median(d.sum(axis=1)):.2f} us")
    print("   " + " | ".join(f"{n} {np.median(d[:, i]):.2f}" for i, n in enumerate(NAMES)))
    full = buf.cpu().numpy().reshape(-1, 16)
    full = full[(full[:, :len(NAMES) + 1] > 0).all(axis=1)]
    if which == "fwd" and (full[:, 12] > 0).all():
        print(f"   pool detail: warp0 loop {np.median(full[:, 12] - full[:, 2]) / 1e3:.2f} | sync {np.median(full[:, 13] - full[:, 12]) / 1e3:.2f} | publish {np.median(full[:, 3] - full[:, 13]) / 1e3:.2f}")
