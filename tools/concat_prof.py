"""Per-kernel CUDA-event durations of the fused concat forward (library profile records). usage: python tools/concat_prof.py [batch]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mga_yolo_b200 import MaskGuidedCBAM, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
lib = _lib.load()
for (Cc, H, W) in [(256, 80, 80), (512, 40, 40), (512, 20, 20)]:
    torch.manual_seed(Cc)
    mod = MaskGuidedCBAM(Cc, sam_cam_fusion="concat", mga_pyramid_fusion="multiply").to(dev)
    x = torch.randn(B, Cc, H, W, device=dev).bfloat16()
    m = torch.randn(B, 1, H, W, device=dev)
    with torch.no_grad():
        for _ in range(3):
            mod([x, m])
        torch.cuda.synchronize()
        lib.mga_profile_enable(1)
        for _ in range(3):
            mod([x, m])
        torch.cuda.synchronize()
    name, val = C.c_char_p(), C.c_float()
    agg = {}
    for i in range(lib.mga_profile_count()):
        lib.mga_profile_read(i, C.byref(name), C.byref(val))
        agg.setdefault(name.value.decode(), []).append(val.value)
    lib.mga_profile_enable(0)
    flops = 2.0 * B * H * W * 2 * Cc * Cc
    print(f"C={Cc} {H}x{W} B={B}: " + " | ".join(f"{k} {sum(v)/len(v)*1e3:.1f} us" for k, v in agg.items()) +
          f" | concat_fwd = {flops / (sum(agg['concat_fwd'])/len(agg['concat_fwd'])) / 1e9:.0f} TFLOP/s, {2*B*Cc*H*W*2/(sum(agg['concat_fwd'])/len(agg['concat_fwd']))/1e6:.0f} GB/s")
