"""Pin the CPU oracle against fixtures produced by the reference itself
(oracle/gen_golden.py: MaskCBAM forward + autograd, MaskUtils/cv2 downsample)."""
import numpy as np
import pytest
import torch

from oracle import cbam_oracle as co
from oracle import mask_oracle as mo
from tests._golden import CBAM_CASES, GOLDEN, PARAM_KEYS, load_cbam, rel_err, t


def _run(rec, dtype, suffix):
    p = co.CbamParams(*(t(rec["p." + k], dtype) for k in PARAM_KEYS))
    x = t(rec["x"], dtype)
    mask = t(rec["mask"], dtype) if bool(rec["has_mask"]) else None
    out, sv = co.cbam_forward(x, mask, p, use_sigmoid_mask=rec["use_sigmoid_mask"], gate_clamp=rec["gated"],
                              feature_dtype=dtype)
    grads = co.cbam_backward(t(rec["g"], dtype), p, sv)
    return out, grads


@pytest.mark.parametrize("tag", CBAM_CASES)
@pytest.mark.parametrize("dtype,suffix,tol", [(torch.float64, "_f64", 1e-12), (torch.float32, "", 2e-6)])
def test_cbam_oracle_matches_reference(tag, dtype, suffix, tol):
    rec = load_cbam(tag)
    if rec["gated"]:
        # ProbMaskGater casts the mask to fp32 (probmaskgater.py:77), so even the reference's
        # fp64 run carries an fp32 sigmoid of the mask: fp32-level agreement is the ceiling.
        tol = max(tol, 5e-7)
    out, grads = _run(rec, dtype, suffix)
    assert rel_err(out, t(rec["out" + suffix])) <= tol
    assert rel_err(grads["dx"], t(rec["dx" + suffix])) <= tol
    if bool(rec["has_mask"]):
        assert grads["dmask"].shape == rec["dmask" + suffix].shape
        assert rel_err(grads["dmask"], t(rec["dmask" + suffix])) <= tol
    for k in PARAM_KEYS:
        ref = t(rec["d." + k + suffix])
        # fp32 parameter gradients are long sums: grade against the fp64 run of the reference
        ref64 = t(rec["d." + k + "_f64"])
        assert grads[k].shape == ref.shape, k
        assert rel_err(grads[k], ref64) <= max(tol, 5e-6 if dtype == torch.float32 else tol), k


def test_closed_form_backward_equals_autograd_all_modes():
    torch.manual_seed(0)
    B, C, H, W = 3, 16, 6, 5
    for scf in co.SAM_CAM_MODES:
        for pyr in co.PYRAMID_MODES:
            p = co.default_params(C, r=4, seed=3, beta=0.2, dtype=torch.float64)
            leaves = [v.clone().requires_grad_(True) for v in (p.w1, p.b1, p.w2, p.b2, p.wsam, p.beta)]
            pp = co.CbamParams(*leaves)
            x = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
            mk = torch.randn(B, 1, H, W, dtype=torch.float64, requires_grad=True)
            g = torch.randn(B, C, H, W, dtype=torch.float64)
            out, sv = co.cbam_forward(x, mk, pp, sam_cam_fusion=scf, mga_pyramid_fusion=pyr)
            out.backward(g)
            got = co.cbam_backward(g, p, sv)
            assert rel_err(got["dx"], x.grad) < 1e-12
            assert rel_err(got["dmask"], mk.grad) < 1e-12
            for k, leaf in zip(PARAM_KEYS, leaves):
                assert rel_err(got[k].detach(), leaf.grad) < 1e-12, (scf, pyr, k)


def test_mask_oracle_bit_exact_vs_cv2_goldens():
    z = np.load(GOLDEN / "mask_downsample.npz")
    sizes, strides = z["sizes"], z["strides"]
    checked = 0
    for si in range(len(sizes)):
        src = z[f"src{si}"]
        for s in strides:
            for method, bridge in (("nearest", "1"), ("area", "1"), ("area", "0"), ("maxpool", "1"),
                                   ("skeleton_bresenham", "1"), ("skeleton_bresenham", "0")):
                ref = z[f"bin{si}_{s}_{method}_{bridge}"]
                got = mo.downsample_mask(src, int(s), method, bridge == "1")
                assert got.dtype == ref.dtype and got.shape == ref.shape
                assert np.array_equal(got, ref), (tuple(sizes[si]), int(s), method, bridge)
                checked += 1
            for method in ("avgpool", "nearest", "area"):
                ref = z[f"prob{si}_{s}_{method}"]
                got = mo.downsample_mask_prob(src, int(s), method)
                assert got.dtype == np.float32 and got.shape == ref.shape
                assert np.array_equal(got, ref), (tuple(sizes[si]), int(s), method)
                checked += 1
    assert checked == len(sizes) * len(strides) * 9


@pytest.mark.parametrize("tag", ["p3", "p4", "p5"])
def test_oracle_matches_reference_on_full_size_neck_shapes(tag):
    """oracle/cbam_oracle.py against outputs of the reference itself on the YOLOv8n P3/P4/P5 shapes (batch 2; inputs regenerated
    from seeds, reference outputs stored sampled by oracle/gen_golden_large.py)."""
    import numpy as np

    from tests._golden import GOLDEN, LARGE_CASES, large_inputs

    z = np.load(GOLDEN / f"cbam_large_{tag}.npz")
    x, mask, g, beta, seed, idx = large_inputs(tag)
    p = co.CbamParams.from_state_dict({k[2:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith("p.")})
    out, sv = co.cbam_forward(x.double(), mask.double(), p)
    gr = co.cbam_backward(g.double(), p, sv)
    assert rel_err(out.reshape(-1)[idx], t(z["out_f64"])) * float(np.abs(z["out_f64"]).max()) / float(z["out_absmax_f64"]) <= 1e-12
    assert rel_err(gr["dx"].reshape(-1)[idx], t(z["dx_f64"])) <= 1e-11
    assert rel_err(gr["dmask"], t(z["dmask_f64"])) <= 1e-11
    for k in PARAM_KEYS:
        assert rel_err(gr[k], t(z["d." + k + "_f64"])) <= 1e-10, k
