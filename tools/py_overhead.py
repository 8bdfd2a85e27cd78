"""Host-side cost of one MaskGuidedCBAM call (tiny shape: the GPU work is negligible)."""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, ".")
from mga_yolo_b200 import MaskGuidedCBAM
dev = torch.device("cuda:0")
m = MaskGuidedCBAM(64).to(dev)
x = torch.randn(1, 64, 8, 8, device=dev)
k = torch.randn(1, 1, 8, 8, device=dev)
def train_step():
    xi = x.clone().requires_grad_(True)
    out = m([xi, k])
    out.sum().backward()
def infer():
    with torch.no_grad():
        m([x, k])
for f in (train_step, infer):
    for _ in range(20): f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300): f()
    torch.cuda.synchronize()
    print(f.__name__, f"{(time.perf_counter()-t0)/300*1e6:.1f} us per call")
pr = cProfile.Profile(); pr.enable()
for _ in range(300): infer()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
pr = cProfile.Profile(); pr.enable()
for _ in range(300): train_step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
