"""GPU parity of the components either side of the block, through nn.Module -> torch op -> C ABI:
MaskECA, the MGAMaskHead 3x3 tail, train-mode ProbMaskGater sampling (Philox noise contract) and the zero-pad collate.
Fixtures come from running the reference (oracle/gen_golden_next.py); the oracles are pinned to them in test_oracle_golden.py.
Tolerances as for the block: fp32 1e-5 (parameter gradients against the fp64 run), 16-bit 1e-2."""
import numpy as np
import pytest
import torch

from oracle import eca_oracle as eo
from oracle import next_oracle as no
from tests._golden import GOLDEN, rel_err, t

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run_eca(mod, x, mask, g):
    x = x.clone().requires_grad_(True)
    mask = None if mask is None else mask.clone().requires_grad_(True)
    for p in mod.parameters():
        p.grad = None
    out = mod(x if mask is None else [x, mask])
    out.backward(g)
    return out.detach(), x.grad, None if mask is None else mask.grad, mod.conv1d.weight.grad.clone(), mod.beta.grad.clone()


@pytest.mark.parametrize("tag", ["basic", "beta", "edge", "nomask", "raw3d"])
def test_eca_golden_cases_match_reference(tag):
    from mga_yolo_b200 import MaskECA

    z = np.load(GOLDEN / f"eca_{tag}.npz")
    C, use_sig = (int(v) for v in z["cfg"])
    mod = MaskECA(C, use_sigmoid_mask=bool(use_sig))
    mod.load_state_dict({"conv1d.weight": t(z["w1d"]), "beta": t(z["beta"])}, strict=True)
    mod.to(DEV)
    mask = t(z["mask"]).to(DEV) if bool(z["has_mask"]) else None
    out, dx, dmask, dw, dbeta = _run_eca(mod, t(z["x"]).to(DEV), mask, t(z["g"]).to(DEV))
    assert rel_err(out.cpu(), t(z["out"])) <= 1e-5
    assert rel_err(dx.cpu(), t(z["dx"])) <= 1e-5
    if mask is not None:
        assert dmask.shape == mask.shape and rel_err(dmask.cpu(), t(z["dmask"])) <= 1e-5
    assert rel_err(dw.cpu(), t(z["d.conv1d.weight_f64"])) <= 2e-5
    assert rel_err(dbeta.cpu(), t(z["d.beta_f64"])) <= 2e-5


@pytest.mark.parametrize("shape", [(4, 64, 80, 80), (2, 256, 20, 20), (2, 96, 17, 13), (1, 8, 8, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_eca_matches_oracle_at_neck_shapes(shape, dtype):
    from mga_yolo_b200 import MaskECA

    B, C, H, W = shape
    gen = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, W, generator=gen).to(dtype)
    mask = torch.randn(B, 1, H, W, generator=gen)
    mask[0] = -20.0
    g = torch.randn(B, C, H, W, generator=gen).to(dtype)
    torch.manual_seed(C)
    mod = MaskECA(C)
    with torch.no_grad():
        mod.beta.fill_(0.3)
    w1d, beta = mod.conv1d.weight.detach().clone().double().reshape(-1), mod.beta.detach().clone().double()
    mod.to(DEV)
    out, dx, dmask, dw, dbeta = _run_eca(mod, x.to(DEV), mask.to(DEV), g.to(DEV))
    ref_out, sv = eo.eca_forward(x.double(), mask.double(), w1d, beta)
    ref = eo.eca_backward(g.double(), sv)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert out.dtype == dtype and rel_err(out.float().cpu(), ref_out) <= tol
    assert rel_err(dx.float().cpu(), ref["dx"]) <= tol
    assert rel_err(dmask.cpu(), ref["dmask"]) <= (1e-5 if dtype == torch.float32 else 1e-2)
    ptol = 5e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(dw.cpu(), ref["conv1d.weight"]) <= ptol
    assert rel_err(dbeta.cpu(), ref["beta"]) <= max(ptol, 4e-5)


def test_eca_module_contract():
    from mga_yolo_b200 import MaskECA
    from mga_yolo_b200.eca import eca_kernel_size

    mod = MaskECA(128).to(DEV)
    assert set(mod.state_dict()) == {"conv1d.weight", "beta"} and mod.conv1d.weight.shape[-1] == eca_kernel_size(128) == eo.eca_kernel_size(128)
    assert abs(float(mod.alpha.detach()) - 0.6931471805599453) < 1e-6
    x = torch.randn(2, 128, 8, 8, device=DEV)
    assert mod(x).shape == x.shape and mod([x, torch.randn(2, 8, 8, device=DEV)]).shape == x.shape  # Tensor / [feat, 3-D mask]
    with pytest.raises(RuntimeError):
        mod(x.cpu())
    with pytest.raises(RuntimeError):
        mod([x, torch.randn(2, 1, 4, 4, device=DEV)])


# ---------------------------------------------------------------- MaskSPADE (feature side in CUDA: csrc/spade_ops.cu)
SPADE_KEYS = ("shared.0.weight", "shared.0.bias", "conv_gamma.weight", "conv_gamma.bias", "conv_beta.weight", "conv_beta.bias")


@pytest.mark.parametrize("tag", ["basic", "nomask", "raw3d", "odd", "resize", "p5"])
def test_spade_golden_cases_match_reference(tag, monkeypatch):
    """Module mirror with the reference's state_dict: out, dx, dmask and the six parameter gradients against the reference run."""
    from mga_yolo_b200 import MaskSPADE

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    z = np.load(GOLDEN / f"spade_{tag}.npz")
    C, hidden, use_sig = (int(v) for v in z["cfg"])
    mod = MaskSPADE(C, hidden=hidden, use_sigmoid_mask=bool(use_sig))
    mod.load_state_dict({k: t(z["p." + k]) for k in SPADE_KEYS}, strict=True)
    mod.to(DEV)
    x = t(z["x"]).to(DEV).requires_grad_(True)
    mask = t(z["mask"]).to(DEV).requires_grad_(True) if bool(z["has_mask"]) else None
    out = mod(x if mask is None else [x, mask])
    out.backward(t(z["g"]).to(DEV))
    assert rel_err(out.detach().cpu(), t(z["out"])) <= 1e-5
    assert rel_err(x.grad.cpu(), t(z["dx"])) <= 1e-5
    if mask is not None:
        assert mask.grad.shape == mask.shape and rel_err(mask.grad.cpu(), t(z["dmask_f64"])) <= 2e-5
        for k in SPADE_KEYS:
            assert rel_err(dict(mod.named_parameters())[k].grad.cpu(), t(z["d." + k + "_f64"])) <= 2e-5, k


@pytest.mark.parametrize("tag", ["basic", "p5", "odd"])
def test_spade_op_given_the_reference_gamma_beta(tag):
    """The CUDA op alone, fed the reference block's own gamma / beta: y, dx, d gamma, d beta (= g) against the fp64 oracle."""
    from mga_yolo_b200 import next_ops
    from oracle import spade_oracle as so

    z = np.load(GOLDEN / f"spade_{tag}.npz")
    x, g, gamma, beta = (t(z[k]).to(DEV) for k in ("x", "g", "gamma", "beta"))
    x.requires_grad_(True), gamma.requires_grad_(True), beta.requires_grad_(True)
    y = next_ops.spade_modulate(x, gamma, beta, 1e-6)
    y.backward(g)
    ref_y, sv = so.modulate_forward(t(z["x"]).double(), t(z["gamma"]).double(), t(z["beta"]).double(), 1e-6)
    ref = so.modulate_backward(t(z["g"]).double(), sv)
    assert rel_err(y.detach().cpu(), ref_y) <= 1e-5 and rel_err(y.detach().cpu(), t(z["out"])) <= 1e-5
    assert rel_err(x.grad.cpu(), ref["dx"]) <= 1e-5
    assert rel_err(gamma.grad.cpu(), ref["dgamma"]) <= 1e-5
    assert torch.equal(beta.grad, g)


@pytest.mark.parametrize("shape", [(2, 64, 80, 80), (2, 256, 20, 20), (2, 128, 40, 40), (1, 8, 128, 128), (1, 4, 150, 200), (2, 6, 9, 7)])
@pytest.mark.parametrize("dtype,mod_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16),
                                             (torch.bfloat16, torch.float32)])
def test_spade_matches_oracle_at_neck_shapes(shape, dtype, mod_dtype):
    """Every launch form: 64 / 128 / 256-thread CTAs, rows staged in shared memory and re-read through L2 (128 x 128: backward only;
    150 x 200: both directions), the scalar path (odd row length), 16-bit features with 16-bit and with fp32 gamma / beta."""
    from mga_yolo_b200 import next_ops
    from oracle import spade_oracle as so

    B, C, H, W = shape
    gen = torch.Generator().manual_seed(C + H)
    x = (torch.randn(B, C, H, W, generator=gen) * 1.5 + 0.3).to(dtype)
    gamma = (1.0 + 0.5 * torch.randn(B, C, H, W, generator=gen)).to(mod_dtype)
    beta = (0.5 * torch.randn(B, C, H, W, generator=gen)).to(mod_dtype)
    g = torch.randn(B, C, H, W, generator=gen).to(dtype)
    xd, gd, bd = x.to(DEV).requires_grad_(True), gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    y = next_ops.spade_modulate(xd, gd, bd, 1e-6)
    y.backward(g.to(DEV))
    ref_y, sv = so.modulate_forward(x.double(), gamma.double(), beta.double(), 1e-6)
    ref = so.modulate_backward(g.double(), sv)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert y.dtype == dtype and rel_err(y.detach().float().cpu(), ref_y) <= tol
    assert xd.grad.dtype == dtype and rel_err(xd.grad.float().cpu(), ref["dx"]) <= tol
    assert gd.grad.dtype == mod_dtype and rel_err(gd.grad.float().cpu(), ref["dgamma"]) <= tol
    assert bd.grad.dtype == mod_dtype and rel_err(bd.grad.float().cpu(), g.double()) <= (0.0 if mod_dtype == dtype else 1e-6)
    # mask-less branch: plain instance norm and its backward
    xn = x.to(DEV).requires_grad_(True)
    yn = next_ops.spade_modulate(xn, None, None, 1e-6)
    yn.backward(g.to(DEV))
    ref_n, svn = so.modulate_forward(x.double(), None, None, 1e-6)
    assert rel_err(yn.detach().float().cpu(), ref_n) <= tol
    assert rel_err(xn.grad.float().cpu(), so.modulate_backward(g.double(), svn)["dx"]) <= tol


def test_spade_module_under_autocast_and_errors():
    from mga_yolo_b200 import MaskSPADE, next_ops

    torch.manual_seed(0)
    mod = MaskSPADE(64, hidden=16).to(DEV)
    x = torch.randn(2, 64, 40, 40, device=DEV, dtype=torch.float16)
    mask = torch.randn(2, 1, 40, 40, device=DEV)
    with torch.autocast("cuda", dtype=torch.float16):  # the reference's AMP training path: fp16 features, fp16 gamma / beta from the convs
        y = mod([x, mask])
    assert y.dtype == torch.float16 and y.shape == x.shape and torch.isfinite(y).all()
    with pytest.raises(RuntimeError, match="does not match"):
        next_ops.spade_modulate(x, torch.zeros(2, 64, 20, 20, device=DEV, dtype=torch.float16), torch.zeros(2, 64, 20, 20, device=DEV, dtype=torch.float16))
    with pytest.raises(RuntimeError):
        next_ops.spade_modulate(x, torch.zeros_like(x), None)


@pytest.mark.parametrize("tag", ["p3", "odd"])
def test_head_tail_matches_reference(tag):
    from mga_yolo_b200 import next_ops

    z = np.load(GOLDEN / f"head_{tag}.npz")
    feat = t(z["feat"]).to(DEV).requires_grad_(True)
    w = t(z["w"]).to(DEV).requires_grad_(True)
    b = t(z["b"]).to(DEV).requires_grad_(True)
    out = next_ops.head_tail(feat, w, b)
    out.backward(t(z["g"]).to(DEV))
    assert rel_err(out.detach().cpu(), t(z["out"])) <= 1e-5
    assert rel_err(feat.grad.cpu(), t(z["dfeat"])) <= 1e-5
    assert rel_err(w.grad.cpu(), t(z["dw_f64"])) <= 2e-5
    assert rel_err(b.grad.cpu(), t(z["db_f64"])) <= 2e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mask_head_module_feeds_the_block(dtype, monkeypatch):
    """MGAMaskHead (tail in the CUDA library) -> MaskGuidedCBAM -> loss: logits, and the gradient that flows back through dmask into the
    head's parameters, against the same modules run with the library convolution for the tail."""
    from mga_yolo_b200 import MGAMaskHead, MaskGuidedCBAM

    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)  # the comparator (library convolution) must be real fp32
    torch.manual_seed(0)
    C, hidden, B, H, W = 64, 32, 2, 40, 40
    head = MGAMaskHead(C, hidden).to(DEV).to(dtype)
    blk = MaskGuidedCBAM(C).to(DEV)
    assert set(head.state_dict()) == {"proj.0.weight", "proj.1.weight", "proj.1.bias", "proj.1.running_mean", "proj.1.running_var",
                                      "proj.1.num_batches_tracked", "head.weight", "head.bias"}
    x = torch.randn(B, C, H, W, device=DEV, dtype=dtype)
    head.eval()

    def step(use_kernel):
        for p in head.parameters():
            p.grad = None
        xi = x.clone().requires_grad_(True)
        feat = head.proj(xi)
        logits = head(xi) if use_kernel else head.head(feat).float()
        out = blk([xi, logits])
        out.float().square().mean().backward()
        return logits.detach().float(), xi.grad.float(), head.head.weight.grad.float().clone(), head.proj[0].weight.grad.float().clone()

    got, ref = step(True), step(False)
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    for a, r in zip(got, ref):
        assert rel_err(a.cpu(), r.cpu()) <= tol


@pytest.mark.parametrize("tag", ["gumbel", "gumbel_pmin", "hard_st"])
def test_gate_sampling_reproduces_reference_given_its_uniforms(tag):
    from mga_yolo_b200 import next_ops

    z = np.load(GOLDEN / f"gate_{tag}.npz")
    tau, p_min, thr = (float(v) for v in z["cfg"])
    p = t(z["p"]).to(DEV).requires_grad_(True)
    noise = torch.stack([t(z["u0"]), t(z["u1"])]).to(DEV)
    out = next_ops.gate_sample(p, str(z["mode"]), tau=tau, p_min=p_min, threshold=thr, noise=noise)
    out.backward(t(z["g"]).to(DEV))
    if str(z["mode"]) == "hard_st":  # a soft value within rounding of the threshold may flip
        assert (out.detach().cpu() != t(z["out"])).float().mean().item() <= 0.005
    else:
        assert rel_err(out.detach().cpu(), t(z["out"])) <= 1e-5
    assert rel_err(p.grad.cpu(), t(z["dp"])) <= 1e-4


def test_gate_noise_contract_is_philox_keyed_by_seed_and_offset():
    from mga_yolo_b200 import _lib, next_ops
    import ctypes as C

    lib = _lib.load()
    n = 4096
    p = torch.full((n,), 0.3, device=DEV)
    out, soft, noise = torch.empty_like(p), torch.empty_like(p), torch.empty(2 * n, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.mga_gate_sample_forward(p.data_ptr(), None, out.data_ptr(), soft.data_ptr(), noise.data_ptr(), n, 0, 1.0, 0.0, 0.5, 1234, 7, st), "gate")
    u1, u2 = no.gate_uniforms(n, seed=1234, offset=7)
    got = noise.cpu().numpy()
    assert np.array_equal(got[:n], u1) and np.array_equal(got[n:], u2)  # bit-exact stream
    ref_out, _ = no.gate_forward(p.cpu(), torch.from_numpy(u1), torch.from_numpy(u2), mode="gumbel", tau=1.0)
    assert rel_err(out.cpu(), ref_out) <= 1e-5
    a = next_ops.gate_sample(p, "gumbel", seed=5, offset=0)
    assert torch.equal(a, next_ops.gate_sample(p, "gumbel", seed=5, offset=0))          # same (seed, offset): same sample
    assert not torch.equal(a, next_ops.gate_sample(p, "gumbel", seed=5, offset=1))      # the caller advances the offset per call
    b = next_ops.gate_sample(torch.full((200000,), 0.3, device=DEV), "bernoulli_detach", seed=9, offset=0)
    assert set(b.unique().tolist()) <= {0.0, 1.0} and abs(b.mean().item() - 0.3) < 0.01  # Bernoulli(p)
    g = next_ops.gate_sample(torch.full((200000,), 0.3, device=DEV), "gumbel", seed=9, offset=3)
    assert abs(g.median().item() - 0.3) < 0.01  # logistic noise is symmetric: the median of sigmoid(logit(p) + noise) is p


def test_train_mode_gate_inside_the_block(monkeypatch):
    from mga_yolo_b200 import MaskGuidedCBAM

    monkeypatch.setenv("MGA_PROB_MODE", "1")
    monkeypatch.setenv("MGA_PROB_APPROACH", "gumbel")
    torch.manual_seed(0)
    blk = MaskGuidedCBAM(32).to(DEV).train()
    x = torch.randn(2, 32, 16, 16, device=DEV, requires_grad=True)
    m = torch.rand(2, 1, 16, 16, device=DEV, requires_grad=True)
    o1 = blk([x, m])
    o1.sum().backward()
    assert torch.isfinite(o1).all() and m.grad is not None and torch.isfinite(m.grad).all() and m.grad.abs().sum() > 0
    assert blk.gater.calls == 1 and not torch.equal(o1, blk([x, m]))  # a fresh noise stream offset per forward


def test_collate_matches_reference_rule():
    from mga_yolo_b200 import MaskUtils

    rng = np.random.default_rng(0)
    for dt in (np.float32, np.uint8):
        per = [(rng.random((h, w)) > 0.5).astype(dt) if dt == np.uint8 else rng.random((h, w)).astype(dt) for h, w in ((10, 8), (7, 12), (10, 12), (1, 1))]
        got = MaskUtils.collate_masks([torch.from_numpy(a).to(DEV) for a in per])
        assert got.dtype == torch.float32 and np.array_equal(got.cpu().numpy(), no.collate_masks(per))


# ---------------------------------------------------------------- sam_cam_fusion = concat: fused tcgen05 forward
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,pyr", [((2, 128, 20, 20), "multiply"), ((2, 256, 16, 24), "add"), ((1, 128, 80, 80), "multiply"), ((3, 256, 40, 40), "add"),
                                       ((2, 256, 80, 80), "multiply"), ((2, 512, 40, 40), "multiply"), ((2, 512, 20, 20), "multiply")])  # BASELINE configs[3] levels
@pytest.mark.parametrize("bwd", ["tc", "library"])  # backward: tcgen05 kernel (mga_cbam_concat_backward_dx) / library GEMM + elementwise kernel
def test_concat_fused_forward_and_backward_match_the_library_composition(shape, pyr, dtype, bwd, monkeypatch):
    """The fused forward (csrc/cbam_concat.cuh: TMA + tcgen05.mma + TMEM epilogue) and its closed-form backward against the same module
    run as gates op + torch.cat + F.conv2d + autograd (MGA_CONCAT_LIBRARY=1) and against the in-repo fp64 oracle.
    Both modes are build-side definitions: 'oracle: in-repo PyTorch composition; reference parity unpinned'."""
    from mga_yolo_b200 import MaskGuidedCBAM
    from oracle import cbam_oracle as co
    from tests._golden import PARAM_KEYS

    B, C, H, W = shape
    monkeypatch.setenv("MGA_CONCAT_BWD", bwd)
    gen = torch.Generator().manual_seed(C + H)
    x = (torch.randn(B, C, H, W, generator=gen) * 0.5).to(dtype)
    mask = torch.randn(B, 1, H, W, generator=gen)
    g = torch.randn(B, C, H, W, generator=gen).to(dtype)
    torch.manual_seed(C)
    mod = MaskGuidedCBAM(C, sam_cam_fusion="concat", mga_pyramid_fusion=pyr).to(DEV)
    with torch.no_grad():
        mod.beta.fill_(0.3)
        mod.fuse_sam_cam.bias.uniform_(-0.2, 0.2)

    def run(library):
        if library:
            monkeypatch.setenv("MGA_CONCAT_LIBRARY", "1")
        else:
            monkeypatch.delenv("MGA_CONCAT_LIBRARY", raising=False)
        for p in mod.parameters():
            p.grad = None
        xi = x.to(DEV).clone().requires_grad_(True)
        mi = mask.to(DEV).clone().requires_grad_(True)
        out = mod([xi, mi])
        out.backward(g.to(DEV))
        return [out.detach().float().cpu(), xi.grad.float().cpu(), mi.grad.float().cpu()] + [p.grad.detach().float().cpu().clone() for p in mod.parameters()]

    fused, lib = run(False), run(True)
    names = ["out", "dx", "dmask"] + [n for n, _ in mod.named_parameters()]
    # the library composition rounds every intermediate to 16 bits (its parameter gradients are 16-bit reductions): compare the
    # activations / activation gradients with it, and everything with the fp64 oracle differentiated by autograd on the rounded inputs
    for n, a, b in list(zip(names, fused, lib))[:3]:
        assert a.shape == b.shape and rel_err(a, b) <= 3e-2, (n, rel_err(a, b))
    sd = {k: v.detach().double().cpu().requires_grad_(True) for k, v in mod.state_dict().items()}
    p = co.CbamParams(*(sd[k] for k in PARAM_KEYS))
    x64 = x.double().requires_grad_(True)
    m64 = mask.double().requires_grad_(True)
    ref_out = co.cbam_forward_general(x64, m64, p, sam_cam_fusion="concat", mga_pyramid_fusion=pyr,
                                      fuse_sam_cam=(sd["fuse_sam_cam.weight"], sd["fuse_sam_cam.bias"]))
    ref_out.backward(g.double())
    assert rel_err(fused[0], ref_out.detach()) <= 1e-2
    assert rel_err(fused[1], x64.grad) <= 2e-2 and rel_err(fused[2], m64.grad) <= 2e-2
    for n, a in zip(names[3:], fused[3:]):
        assert rel_err(a, sd[n].grad) <= 2e-2, (n, rel_err(a, sd[n].grad))
