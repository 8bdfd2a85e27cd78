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


# ---------------------------------------------------------------- MaskECA, MGAMaskHead tail, ProbMaskGater (train mode)
ECA_CASES = ("basic", "beta", "edge", "nomask", "raw3d")


@pytest.mark.parametrize("tag", ECA_CASES)
@pytest.mark.parametrize("dtype,suffix,tol", [(torch.float64, "_f64", 1e-12), (torch.float32, "", 2e-6)])
def test_eca_oracle_matches_reference(tag, dtype, suffix, tol):
    from oracle import eca_oracle as eo

    z = np.load(GOLDEN / f"eca_{tag}.npz")
    C, use_sig = (int(v) for v in z["cfg"])
    assert z["w1d"].shape[-1] == eo.eca_kernel_size(C)
    x = t(z["x"], dtype)
    mask = t(z["mask"], dtype) if bool(z["has_mask"]) else None
    out, sv = eo.eca_forward(x, mask, t(z["w1d"], dtype).reshape(-1), t(z["beta"], dtype), use_sigmoid_mask=bool(use_sig))
    gr = eo.eca_backward(t(z["g"], dtype), sv)
    assert rel_err(out, t(z["out" + suffix])) <= tol
    assert rel_err(gr["dx"], t(z["dx" + suffix])) <= tol
    if mask is not None:
        assert gr["dmask"].shape == z["dmask" + suffix].shape
        assert rel_err(gr["dmask"], t(z["dmask" + suffix])) <= tol
    for k in ("conv1d.weight", "beta"):
        assert rel_err(gr[k], t(z["d." + k + "_f64"])) <= max(tol, 2e-5 if dtype == torch.float32 else 0.0), k


@pytest.mark.parametrize("tag", ["p3", "odd"])
@pytest.mark.parametrize("dtype,suffix,tol", [(torch.float64, "_f64", 1e-12), (torch.float32, "", 2e-6)])
def test_head_tail_oracle_matches_reference(tag, dtype, suffix, tol):
    from oracle import next_oracle as no

    z = np.load(GOLDEN / f"head_{tag}.npz")
    feat, w, b, g = (t(z[k], dtype) for k in ("feat", "w", "b", "g"))
    out = no.head_tail_forward(feat, w, b)
    dfeat, dw, db = no.head_tail_backward(feat, w, g)
    assert rel_err(out, t(z["out" + suffix])) <= tol
    assert rel_err(dfeat, t(z["dfeat" + suffix])) <= tol
    assert rel_err(dw, t(z["dw_f64"])) <= max(tol, 2e-5 if dtype == torch.float32 else 0.0)
    assert rel_err(db, t(z["db_f64"])) <= max(tol, 2e-5 if dtype == torch.float32 else 0.0)


@pytest.mark.parametrize("tag", ["gumbel", "gumbel_pmin", "hard_st"])
def test_gate_oracle_matches_reference_given_its_uniform_draws(tag):
    from oracle import next_oracle as no

    z = np.load(GOLDEN / f"gate_{tag}.npz")
    tau, p_min, thr = (float(v) for v in z["cfg"])
    assert int(z["n_u"]) == 2
    out, saved = no.gate_forward(t(z["p"]), t(z["u0"]), t(z["u1"]), mode=str(z["mode"]), tau=tau, p_min=p_min, threshold=thr)
    assert rel_err(out, t(z["out"])) <= 2e-6
    dp = no.gate_backward(t(z["g"]), saved, tau=tau, p_min=p_min)
    assert rel_err(dp, t(z["dp"])) <= 2e-5


def test_philox_known_answers():
    """Random123 known-answer vectors of Philox4x32-10."""
    from oracle import next_oracle as no

    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = no.philox4x32(ctr, key)
        assert tuple(int(v) for v in got) == want
    u1, u2 = no.gate_uniforms(1000, seed=1234, offset=7)
    assert u1.dtype == np.float32 and (u1 > 0).all() and (u1 <= 1).all() and abs(float(u1.mean()) - 0.5) < 0.05 and abs(float(u2.mean()) - 0.5) < 0.05


def test_collate_oracle_matches_reference_rule():
    """dataset.py:149-169: F.pad(t, (0, pad_w, 0, pad_h), 0) then stack."""
    import torch.nn.functional as F

    from oracle import next_oracle as no

    rng = np.random.default_rng(0)
    per = [rng.random((1, h, w)).astype(np.float32) for h, w in ((10, 8), (7, 12), (10, 12))]
    want = torch.stack([F.pad(torch.from_numpy(a), (0, 12 - a.shape[2], 0, 10 - a.shape[1]), value=0.0) for a in per], 0).numpy()
    assert np.array_equal(no.collate_masks([a[0] for a in per]), want)


# ---------------------------------------------------------------- MaskSPADE
SPADE_CASES = ("basic", "nomask", "raw3d", "odd", "resize", "p5")


@pytest.mark.parametrize("tag", SPADE_CASES)
@pytest.mark.parametrize("dtype,suffix,tol", [(torch.float64, "_f64", 1e-12), (torch.float32, "", 5e-6)])
def test_spade_oracle_matches_reference(tag, dtype, suffix, tol):
    from oracle import spade_oracle as so

    z = np.load(GOLDEN / f"spade_{tag}.npz")
    _, _, use_sig = (int(v) for v in z["cfg"])
    params = {k: t(z["p." + k], dtype) for k in so.PARAM_KEYS}
    mask = t(z["mask"], dtype) if bool(z["has_mask"]) else None
    res = so.spade_forward_backward(t(z["x"], dtype), mask, params, t(z["g"], dtype), use_sigmoid_mask=bool(use_sig))
    assert rel_err(res["out"], t(z["out" + suffix])) <= tol
    assert rel_err(res["dx"], t(z["dx" + suffix])) <= tol
    if mask is not None:
        assert res["dmask"].shape == z["dmask" + suffix].shape and rel_err(res["dmask"], t(z["dmask" + suffix])) <= tol
        assert rel_err(res["gamma"], t(z["gamma" + suffix])) <= tol and rel_err(res["beta"], t(z["beta" + suffix])) <= tol
        for k in so.PARAM_KEYS:
            assert rel_err(res[k], t(z["d." + k + "_f64"])) <= max(tol, 2e-5 if dtype == torch.float32 else 0.0), k
