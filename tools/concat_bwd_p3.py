"""ncu target: one tcgen05 concat backward at the cfg4-P3 shape (resident-weight kernel, MODE 1), batch 32."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from mga_yolo_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0")
B, Cc, H, W = 32, 256, 80, 80
S = H * W; dt = torch.bfloat16
x = torch.randn(B, Cc, H, W, device=dev).to(dt); g = torch.randn_like(x)
s = torch.rand(B, Cc, device=dev); a = torch.rand(B, S, device=dev); w = torch.randn(Cc, 2 * Cc, device=dev) * 0.05
bias = torch.zeros(Cc, device=dev); beta = torch.zeros((), device=dev)
nT = (S + 127) // 128
dx = torch.empty_like(x); ga = torch.empty_like(x)
dsp = torch.empty(B, 2 * nT, Cc, device=dev); dbp = torch.empty_like(dsp); dap = torch.empty(B, Cc // 32, S, device=dev); dal = torch.empty(B, nT, Cc // 16, device=dev)
ws = torch.empty(2 * Cc * Cc, dtype=dt, device=dev)
d = _lib.Desc(B, Cc, H, W, 1, 1, _lib.BF16, _lib.F32, _lib.PYRAMID_MULTIPLY, 0.0, 0.0)
for _ in range(3):
    rc = lib.mga_cbam_concat_backward_dx(C.byref(d), x.data_ptr(), g.data_ptr(), s.data_ptr(), a.data_ptr(), w.data_ptr(), bias.data_ptr(), beta.data_ptr(),
                                         dx.data_ptr(), ga.data_ptr(), dsp.data_ptr(), dbp.data_ptr(), dap.data_ptr(), dal.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print("rc", rc)
