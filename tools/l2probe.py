"""Per-kernel CUDA-event times of the split path at several batch sizes (is a re-read of x an L2 hit, and how fast?)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from bench import DT, KERNEL_BYTES, WORKLOADS, LevelPlan  # noqa: E402
from mga_yolo_b200 import _lib  # noqa: E402

lib = _lib.load()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
levels, _, dtname, _ = WORKLOADS[wl]
dev = torch.device("cuda:0")
es = torch.empty((), dtype=DT[dtname]).element_size()
for li, (Cc, H, W) in enumerate(levels[:1] if "--p3" in sys.argv else levels):
    for B in (4, 8, 16, 32, 64):
        flat = torch.zeros(LevelPlan.n_params(Cc), device=dev)
        pl = LevelPlan(None, Cc, H, W, B, DT[dtname], _lib.FORCE_SPLIT, dev, li, flat, 0)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            pl.fwd(st); pl.bwd(st)
        torch.cuda.synchronize()
        lib.mga_profile_enable(1)
        reps = 5
        for _ in range(reps):
            pl.fwd(st); pl.bwd(st)
        torch.cuda.synchronize()
        name, val = C.c_char_p(), C.c_float()
        acc = {}
        n = lib.mga_profile_count()
        per = n // reps
        for i in range(per, n):
            lib.mga_profile_read(i, C.byref(name), C.byref(val))
            acc.setdefault((i % per, name.value.decode()), []).append(val.value)
        lib.mga_profile_enable(0)
        N, BS = B * Cc * H * W, B * H * W
        line = []
        for (j, nm), v in sorted(acc.items()):
            ms = sum(v) / len(v)
            fn = KERNEL_BYTES.get(nm)
            line.append(f"{nm}={ms*1e3:.1f}us" + (f"({fn(N, BS, es)/ms/1e6:.0f}GB/s)" if fn else ""))
        print(f"P{3+li} B={B} x={N*es/1e6:.0f}MB: " + " ".join(line))
