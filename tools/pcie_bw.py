"""Raw pinned-host <-> device copy bandwidth on this box (explains bench.py's e2e numbers): H2D alone, D2H alone, both at once."""
import torch

dev = torch.device("cuda:0")
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.ones(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("both", both, 2 * n)):
    ms = timed(fn)
    print(f"{name}: {ms:.2f} ms for {nbytes >> 20} MiB -> {nbytes / ms / 1e6:.1f} GB/s")
# fresh (non-reused) pinned destination each time, as an allocator would hand out
ms = timed(lambda: torch.empty(n, dtype=torch.uint8).pin_memory().copy_(d_b, non_blocking=True), reps=2)
print(f"d2h into a freshly pinned buffer: {ms:.2f} ms")
