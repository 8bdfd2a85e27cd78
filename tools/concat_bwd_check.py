"""Direct C-ABI check of mga_cbam_concat_backward_dx (tcgen05 backward of the concat mode) against a torch composition; then the same
through the nn.Module on autograd's thread.  usage: python tools/concat_bwd_check.py [512]  (512: one cfg4-P4-shaped call, the ncu target)"""
import ctypes as C, torch, sys
sys.path.insert(0, ".")
from mga_yolo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
SHAPES = [(32, 512, 40, 40)] if len(sys.argv) > 1 and sys.argv[1] == "512" else [(2, 128, 20, 20), (2, 256, 16, 24), (128, 256, 80, 80)]
for (B, Cc, H, W) in SHAPES:
    S = H * W
    dt = torch.bfloat16
    x = torch.randn(B, Cc, H, W, device=dev).to(dt); g = torch.randn_like(x)
    s = torch.rand(B, Cc, device=dev); a = torch.rand(B, S, device=dev); w = torch.randn(Cc, 2 * Cc, device=dev) * 0.05
    bias = torch.zeros(Cc, device=dev); beta = torch.zeros((), device=dev)
    nT = (S + 127) // 128
    dx = torch.empty_like(x); ga = torch.empty_like(x)
    EW = int(__import__('os').getenv('MGA_CC_EPI_WARPS', '8'))
    dsp = torch.empty(B, (EW // 4) * nT, Cc, device=dev); dbp = torch.empty_like(dsp); dap = torch.empty(B, Cc // 32, S, device=dev); dal = torch.empty(B, nT, (Cc // 128) * EW, device=dev)
    ws = torch.empty(2 * Cc * Cc, dtype=dt, device=dev)
    d = _lib.Desc(B, Cc, H, W, 1, 1, _lib.BF16, _lib.F32, _lib.PYRAMID_MULTIPLY, 0.0, 0.0)
    rc = lib.mga_cbam_concat_backward_dx(C.byref(d), x.data_ptr(), g.data_ptr(), s.data_ptr(), a.data_ptr(), w.data_ptr(), bias.data_ptr(), beta.data_ptr(),
                                         dx.data_ptr(), ga.data_ptr(), dsp.data_ptr(), dbp.data_ptr(), dap.data_ptr(), dal.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print((B, Cc, H, W), "rc", rc, lib.mga_last_error() if rc else "")
    if rc == 0:
        xf, gf = x.float().reshape(B, Cc, S), g.float().reshape(B, Cc, S)
        wa, wb = w[:, :Cc].to(dt).float(), w[:, Cc:].to(dt).float()
        U = torch.matmul(wa.t(), gf); V = torch.matmul(wb.t(), gf)
        alpha = torch.nn.functional.softplus(beta)
        ref = alpha * (s[:, :, None] * U + a[:, None, :] * V)
        print("  dx err", float((dx.float().reshape(B, Cc, S) - ref).abs().max() / ref.abs().max()),
              " ga err", float((ga.float().reshape(B, Cc, S) - gf * a[:, None, :]).abs().max()),
              " ds err", float((dsp.sum(1) - alpha * (xf * U).sum(2)).abs().max() / (xf * U).sum(2).abs().max()),
              " da err", float((dap.sum(1) - alpha * (xf * V).sum(1)).abs().max() / (xf * V).sum(1).abs().max()),
              " db err", float((dbp.sum((0, 1)) - alpha * gf.sum((0, 2))).abs().max() / gf.sum((0, 2)).abs().max()))

# the same through the nn.Module (autograd thread)
from mga_yolo_b200 import MaskGuidedCBAM
for (B, Cc, H, W) in [(2, 128, 20, 20)]:
    mod = MaskGuidedCBAM(Cc, sam_cam_fusion="concat", mga_pyramid_fusion="multiply").to(dev)
    x = torch.randn(B, Cc, H, W, device=dev).bfloat16().requires_grad_(True)
    m = torch.randn(B, 1, H, W, device=dev).requires_grad_(True)
    out = mod([x, m])
    try:
        out.backward(torch.randn_like(out))
        print("module backward ok", float(x.grad.float().abs().mean()))
    except Exception as e:
        print("module backward FAILED:", e)
