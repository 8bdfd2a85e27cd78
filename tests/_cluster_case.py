"""Child process of test_gpu_cluster_geometry.py: one forward+backward through the cluster-per-sample kernels with a FORCED
cluster size (MGA_CL_CS_F / MGA_CL_CS_B are read once per process), checked against the oracle.
usage: python -m tests._cluster_case B C H W dtype sam_cam pyramid [default|nomask|rawmask|nograd_mask]"""
import sys

import torch

from oracle import cbam_oracle as co
from tests._golden import PARAM_KEYS, rel_err


def main():
    B, C, H, W = map(int, sys.argv[1:5])
    dtype = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}[sys.argv[5]]
    scf, pyr = sys.argv[6], sys.argv[7]
    variant = sys.argv[8] if len(sys.argv) > 8 else "default"
    from mga_yolo_b200 import MaskGuidedCBAM

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(B * 131 + C + H)
    x = torch.randn(B, C, H, W, generator=gen).to(dtype)
    mask = torch.randn(B, 1, H, W, generator=gen)
    mask[0] = -20.0  # tiny-mask + no-valid-pixel fall-backs
    use_sig = True
    if variant == "rawmask":  # binary {0,1} map, use_sigmoid_mask=False (masked_cbam.py:93-94,110-111,143-144)
        mask = (mask > 0.3).float()
        use_sig = False
    if variant == "nomask":
        mask = None
    g = torch.randn(B, C, H, W, generator=gen).to(dtype)
    p = co.default_params(C, seed=C, beta=0.25)
    mod = MaskGuidedCBAM(C, sam_cam_fusion=scf, mga_pyramid_fusion=pyr, use_sigmoid_mask=use_sig)
    mod.load_state_dict(dict(zip(PARAM_KEYS, (p.w1, p.b1, p.w2, p.b2, p.wsam, p.beta))))
    mod.to(dev)
    xd = x.to(dev).requires_grad_(True)
    want_dmask = variant in ("default", "rawmask")
    md = None if mask is None else mask.to(dev).requires_grad_(want_dmask)
    out = mod(xd if md is None else [xd, md])
    out.backward(g.to(dev))
    ref_out, sv = co.cbam_forward(x.double(), None if mask is None else mask.double(), p.to(torch.float64), sam_cam_fusion=scf,
                                  mga_pyramid_fusion=pyr, feature_dtype=dtype, use_sigmoid_mask=use_sig)
    ref = co.cbam_backward(g.double(), p.to(torch.float64), sv)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    errs = {"out": rel_err(out.detach().float().cpu(), ref_out), "dx": rel_err(xd.grad.float().cpu(), ref["dx"])}
    if want_dmask:
        errs["dmask"] = rel_err(md.grad.float().cpu(), ref["dmask"])
    elif md is not None:
        assert md.grad is None
    ptol = 2e-5 if dtype == torch.float32 else 1e-4
    for k, prm in zip(PARAM_KEYS, mod.parameters()):
        pass
    grads = {n: q.grad.detach().cpu() for n, q in mod.named_parameters()}
    for k in PARAM_KEYS:
        errs["d." + k] = rel_err(grads[k], ref[k])
    # parameter gradients are long cancelling fp32 sums (d beta sums all N elements): same tie-breaker as test_gpu_cbam.py -- when the
    # fp32 ORACLE itself is further than the gate from fp64, the kernel may be up to 4x the fp32 oracle's own error away
    _, sv32 = co.cbam_forward(x.float(), mask, p, sam_cam_fusion=scf, mga_pyramid_fusion=pyr, feature_dtype=dtype, use_sigmoid_mask=use_sig)
    ref32 = co.cbam_backward(g.float(), p, sv32)
    slack = {"d." + k: 4.0 * rel_err(ref32[k], ref[k]) for k in PARAM_KEYS}
    bad = {k: v for k, v in errs.items() if v > (max(ptol, slack[k]) if k in slack else tol)}
    print("ERRS", {k: float(f"{v:.3g}") for k, v in errs.items()})
    if bad:
        print("FAIL", bad)
        return 1
    print("OK")
    return 0


if __name__ == "__main__":
    sys.exit(main())
