"""Run forward(+backward) of ONE pyramid level a few times through the C ABI (ncu target).
usage: python tools/run_level.py <workload> <level index> <fwd|bwd|both> [reps] [--split]"""
import sys

import torch

sys.path.insert(0, ".")
from bench import DT, WORKLOADS, LevelPlan  # noqa: E402
from mga_yolo_b200 import _lib  # noqa: E402

wl, li, which = sys.argv[1], int(sys.argv[2]), sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 3
flags = _lib.FORCE_SPLIT if "--split" in sys.argv else 0
if "--add" in sys.argv:
    flags |= _lib.SAMCAM_ADD
if "--flow" in sys.argv:
    raise SystemExit("the dataflow experiment was removed in round 2")
levels, B, dtname, _ = WORKLOADS[wl]
Cc, H, W = levels[li]
dev = torch.device("cuda:0")
flat = torch.zeros(LevelPlan.n_params(Cc), device=dev)
pl = LevelPlan(None, Cc, H, W, B, DT[dtname], flags, dev, li, flat, 0)
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    if which in ("fwd", "both"):
        pl.fwd(st)
    if which in ("bwd", "both"):
        if which == "bwd" and _ == 0:
            pl.fwd(st)
        pl.bwd(st)
torch.cuda.synchronize()
print("ok", float(pl.out.float().abs().mean()))
