"""Persistent path vs cluster path: parity of forward/backward on the BASELINE shapes and CUDA-event timings (rotating buffers).
usage: python tools/pf_check.py [cfg2|cfg3] [batch]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mga_yolo_b200 import MaskGuidedCBAM  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
levels, dtype = {"cfg2": ([(64, 80, 80), (128, 40, 40), (256, 20, 20)], torch.float32),
                 "cfg3": ([(128, 80, 80), (256, 40, 40), (512, 20, 20)], torch.bfloat16)}[cfg]
dev = torch.device("cuda:0")


def run(mod, x, m, g, no_persist):
    if no_persist:
        os.environ["MGA_NO_PERSIST"] = "1"
    else:
        os.environ.pop("MGA_NO_PERSIST", None)
    x = x.clone().requires_grad_(True)
    m = m.clone().requires_grad_(True)
    for p in mod.parameters():
        p.grad = None
    out = mod([x, m])
    out.backward(g)
    return [out.detach(), x.grad, m.grad] + [p.grad.clone() for p in mod.parameters()]


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def timeit(fn, sets, n=30):
    for i in range(5):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(sets[i % len(sets)])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (C, H, W) in levels:
    torch.manual_seed(C)
    mod = MaskGuidedCBAM(C).to(dev)
    with torch.no_grad():
        mod.beta.fill_(0.3)
    nset = 4 if B * C * H * W * 4 < 300e6 else 2
    sets = [(torch.randn(B, C, H, W, device=dev).to(dtype), torch.randn(B, 1, H, W, device=dev), torch.randn(B, C, H, W, device=dev).to(dtype))
            for _ in range(nset)]
    x, m, g = sets[0]
    m[0] = -20.0
    t0 = time.time()
    ra = run(mod, x, m, g, False)
    torch.cuda.synchronize()
    rb = run(mod, x, m, g, True)
    torch.cuda.synchronize()
    errs = [rel(a, b) for a, b in zip(ra, rb)]
    print(f"C={C} {H}x{W} {dtype} B={B}: persist vs cluster rel err out/dx/dmask/params:", " ".join(f"{e:.1e}" for e in errs), flush=True)
    e = x.element_size()
    N = B * C * H * W
    for name, nop in (("persist", False), ("cluster", True)):
        if nop:
            os.environ["MGA_NO_PERSIST"] = "1"
        else:
            os.environ.pop("MGA_NO_PERSIST", None)

        def fwd(s):
            with torch.no_grad():
                mod([s[0], s[1]])

        xs = [(s[0].clone().requires_grad_(True), s[1].clone().requires_grad_(True), s[2]) for s in sets]

        def fb(s):
            out = mod([s[0], s[1]])
            out.backward(s[2])

        tf = timeit(fwd, sets)
        tfb = timeit(fb, xs)
        print(f"   {name}: fwd(no_save) {tf:7.1f} us = {(2*N+B*H*W)*e/tf/1e3:6.0f} GB/s ({(2*N+B*H*W)*e/tf/1e3/6453.4:.3f}) | fwd+bwd {tfb:7.1f} us = {(5*N+3*B*H*W)*e/tfb/1e3:6.0f} GB/s ({(5*N+3*B*H*W)*e/tfb/1e3/6453.4:.3f})",
              flush=True)
