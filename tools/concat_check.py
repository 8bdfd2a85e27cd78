"""cfg4 (BASELINE configs[3]) concat / multiply: fused tcgen05 forward vs the library composition -- launches and CUDA-event timings.
usage: python tools/concat_check.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mga_yolo_b200 import MaskGuidedCBAM, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
lib = _lib.load()
for (C, H, W) in [(256, 80, 80), (512, 40, 40), (512, 20, 20)]:
    torch.manual_seed(C)
    mod = MaskGuidedCBAM(C, sam_cam_fusion="concat", mga_pyramid_fusion="multiply").to(dev)
    x = torch.randn(B, C, H, W, device=dev).bfloat16()
    m = torch.randn(B, 1, H, W, device=dev)
    g = torch.randn(B, C, H, W, device=dev).bfloat16()
    for name, env in (("fused", ""), ("library", "1")):
        if env:
            os.environ["MGA_CONCAT_LIBRARY"] = env
        else:
            os.environ.pop("MGA_CONCAT_LIBRARY", None)

        def fwd():
            with torch.no_grad():
                return mod([x, m])

        def fb():
            xi = x.detach().requires_grad_(True)
            mi = m.detach().requires_grad_(True)
            mod([xi, mi]).backward(g)
            mod.zero_grad(set_to_none=True)

        res = {}
        for tag, fn in (("fwd", fwd), ("fwd+bwd", fb)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            n0 = lib.mga_launch_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            res[tag] = (a.elapsed_time(b) / 5, (lib.mga_launch_count() - n0) // 5)
        N = B * C * H * W
        flops = 2.0 * B * H * W * 2 * C * C
        print(f"C={C} {H}x{W} B={B} {name:8s}: fwd {res['fwd'][0]*1e3:8.1f} us ({res['fwd'][1]} lib launches) = {2*N*2/res['fwd'][0]/1e6:6.0f} GB/s, "
              f"{flops/res['fwd'][0]/1e9:6.1f} TFLOP/s | fwd+bwd {res['fwd+bwd'][0]*1e3:8.1f} us", flush=True)
