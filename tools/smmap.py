"""Debug: which SM every CTA of the cluster kernels lands on (first wave pairing).  usage: python tools/smmap.py cfg2 fwd|bwd [level]"""
import ctypes as C
import sys
import collections
import numpy as np
import torch

sys.path.insert(0, ".")
from bench import LevelPlan, WORKLOADS, DT  # noqa: E402
from mga_yolo_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.mga_debug_timeline.argtypes = [C.c_void_p]
wl, which = sys.argv[1], sys.argv[2]
li = int(sys.argv[3]) if len(sys.argv) > 3 else 0
levels, B, dtname, _ = WORKLOADS[wl]
Cc, H, W = levels[li]
dev = torch.device("cuda:0")
flat = torch.zeros(LevelPlan.n_params(Cc), device=dev)
pl = LevelPlan(None, Cc, H, W, B, DT[dtname], 0, dev, li, flat, 0)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
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
n = int((t[:, 0] > 0).sum())
t = t[:n]
t0 = t[:, 0].min()
start = (t[:, 0] - t0) / 1e3
smid = t[:, 15]
CS = 8
first = start < 5.0
print(f"{n} CTAs, first wave {int(first.sum())}")
bysm = collections.defaultdict(list)
for i in range(n):
    if first[i]:
        bysm[int(smid[i])].append(i // CS)
pairs = collections.Counter()
for sm, cl in sorted(bysm.items()):
    pairs[tuple(sorted(c & 1 for c in cl))] += 1
print("SMs by (cluster parity) of their first-wave CTAs:", dict(pairs))
print("first 24 SMs:", {sm: bysm[sm] for sm in sorted(bysm)[:24]})
# clusters -> SMs
cl0 = [int(smid[i]) for i in range(8)]
print("cluster 0 on SMs", cl0, " cluster 1 on", [int(smid[i]) for i in range(8, 16)], " cluster 2 on", [int(smid[i]) for i in range(16, 24)])
