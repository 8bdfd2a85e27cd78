"""GPU parity of the CUDA path against THE REFERENCE CLASS ITSELF, run on the same GPU in float64.

oracle/_ref (laid out by oracle/build_ref.py from the reference's own source files; it travels to the GPU box with the
snapshot) holds the unmodified `MaskCBAM` / `MGAMaskHead`.  Here the reference runs on cuda:0 in fp64 with torch autograd --
an oracle that needs no sampling, so the BASELINE configs that round 1 left untested are compared WHOLE:

  * configs[1] YOLOv8n fp32, batch 3 (odd: a short last wave of clusters), every output and all 8 gradients at 1e-5;
  * configs[2] YOLOv8s P3/P4/P5 (128x80x80, 256x40x40, 512x20x20) in bf16: reference fp64 on the bf16-rounded inputs, 1e-2;
  * configs[4] YOLOv8x@1280 P3/P4/P5 (384x160x160, 768x80x80, 768x40x40) in bf16 -- the per-phase path incl. `bwd_partsum`;
  * the producer/consumer pair of the reference graph (yolov8_cbam.yaml:67-72): MGAMaskHead -> logits -> MaskCBAM, with the
    mask gradient flowing back into the head's parameters.
Tolerances (BASELINE.json north_star): fp32 max|d|/max|ref| <= 1e-5, bf16 <= 1e-2.
"""
import pytest
import torch

from tests._golden import PARAM_KEYS, rel_err

pytestmark = pytest.mark.gpu
FP32_TOL, LOWP_TOL = 1e-5, 1e-2


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref

    if not build_ref.available():
        pytest.skip("oracle/_ref is not built (python oracle/build_ref.py in the authoring container)")
    return build_ref.load()


def _pair(ref, C, dev, beta=0.35, seed=0):
    """(our block, the reference block in fp64) with identical parameter values."""
    from mga_yolo_b200 import MaskGuidedCBAM

    torch.manual_seed(seed + C)
    ours = MaskGuidedCBAM(C)
    with torch.no_grad():
        ours.beta.fill_(beta)
    theirs = ref.MaskCBAM(C)
    theirs.load_state_dict(ours.state_dict())
    return ours.to(dev), theirs.to(dev).double()


def _inputs(shape, dtype, dev, seed):
    B, C, H, W = shape
    gen = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=gen, device=dev).to(dtype)
    mask = (torch.randn(B, 1, H, W, generator=gen, device=dev) * 2.0).to(dtype)
    mask[-1, :, : H // 4] = -9.0  # a masked-out band
    g = torch.randn(B, C, H, W, generator=gen, device=dev).to(dtype)
    return x, mask, g


def _run(mod, x, mask, g):
    xi = x.detach().clone().requires_grad_(True)
    mi = mask.detach().clone().requires_grad_(True)
    mod.zero_grad(set_to_none=True)
    out = mod([xi, mi])
    out.backward(g)
    grads = {k: p.grad.detach().clone() for k, p in mod.named_parameters()}
    return out.detach(), xi.grad, mi.grad, grads


def _compare(ref, shape, dtype, tol, grad_tol, monkeypatch, split):
    dev = torch.device("cuda:0")
    if split:
        monkeypatch.setenv("MGA_FORCE_SPLIT", "1")
    else:
        monkeypatch.delenv("MGA_FORCE_SPLIT", raising=False)
    ours, theirs = _pair(ref, shape[1], dev)
    x, mask, g = _inputs(shape, dtype, dev, seed=shape[1] + shape[2])
    out, dx, dmask, grads = _run(ours, x, mask, g)
    rout, rdx, rdmask, rgrads = _run(theirs, x.double(), mask.double(), g.double())
    assert out.dtype == dtype and dx.dtype == dtype
    assert rel_err(out.float(), rout) <= tol
    assert rel_err(dx.float(), rdx) <= tol
    assert rel_err(dmask.float(), rdmask) <= tol
    for k in PARAM_KEYS:
        assert rel_err(grads[k], rgrads[k]) <= grad_tol, k


@pytest.mark.parametrize("split", [False, True], ids=["cluster", "split"])
@pytest.mark.parametrize("shape", [(3, 64, 80, 80), (3, 128, 40, 40), (3, 256, 20, 20)], ids=["P3", "P4", "P5"])
def test_yolov8n_fp32_matches_reference_class_everywhere(ref, shape, split, monkeypatch):
    _compare(ref, shape, torch.float32, FP32_TOL, FP32_TOL, monkeypatch, split)


@pytest.mark.parametrize("split", [False, True], ids=["cluster", "split"])
@pytest.mark.parametrize("shape", [(2, 128, 80, 80), (2, 256, 40, 40), (2, 512, 20, 20)], ids=["P3", "P4", "P5"])
def test_yolov8s_bf16_shapes_match_reference_fp64_on_rounded_inputs(ref, shape, split, monkeypatch):
    """BASELINE configs[2] shapes.  Parameter gradients are fp32 sums of exact products of the rounded inputs: 1e-4."""
    _compare(ref, shape, torch.bfloat16, LOWP_TOL, 1e-3, monkeypatch, split)


@pytest.mark.parametrize("shape", [(1, 384, 160, 160), (1, 768, 80, 80), (2, 768, 40, 40)], ids=["P3", "P4", "P5"])
def test_yolov8x_1280_bf16_shapes_match_reference(ref, shape, monkeypatch):
    """BASELINE configs[4] shapes: 19.7 MB / 9.8 MB samples take the per-phase kernels (incl. bwd_partsum), P5 the clusters."""
    from mga_yolo_b200 import ops

    assert ops.plan(shape, torch.bfloat16, backward=True)["path"] == ("cluster" if shape[2] == 40 else "per_phase")
    _compare(ref, shape, torch.bfloat16, LOWP_TOL, 1e-3, monkeypatch, split=False)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_mask_head_feeds_the_block_and_receives_its_mask_gradient(ref, dtype):
    """The reference graph around the hot path (yolov8_cbam.yaml:67-72): logits = MGAMaskHead(feat); refined = MaskCBAM([feat, logits]).
    The CUDA block replaces MaskCBAM only; d(loss)/d(logits) it returns must train the head exactly as the reference's autograd does."""
    dev = torch.device("cuda:0")
    C, H, W = 64, 40, 40
    ours, theirs = _pair(ref, C, dev, beta=0.2, seed=3)
    torch.manual_seed(9)
    head = ref.MGAMaskHead(C, C // 4).to(dev)  # hidden = C/4 as parse_model scales it (tasks.py:1724-1731)
    head64 = ref.MGAMaskHead(C, C // 4).to(dev).double()
    head64.load_state_dict(head.state_dict())
    head.eval(), head64.eval()  # BatchNorm in inference mode: no cross-sample statistics in the comparison
    gen = torch.Generator(device=dev).manual_seed(4)
    feat = torch.randn(2, C, H, W, generator=gen, device=dev).to(dtype)
    g = torch.randn(2, C, H, W, generator=gen, device=dev).to(dtype)

    def step(block, hd, f, gg, cast):
        fi = f.detach().clone().requires_grad_(True)
        hd.zero_grad(set_to_none=True)
        block.zero_grad(set_to_none=True)
        logits = hd(cast(fi))
        logits = logits if isinstance(logits, torch.Tensor) else logits[0]
        out = block([fi, logits])  # fp32 logits beside bf16 features: the op takes the mask in its own dtype
        out.backward(gg)
        return out.detach(), fi.grad, {k: p.grad.detach().clone() for k, p in hd.named_parameters() if p.grad is not None}

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False  # the head's convolutions are library kernels: keep them fp32-exact for the comparison
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out, dfeat, hg = step(ours, head, feat, g, lambda t_: t_.float())
        rout, rdfeat, rhg = step(theirs, head64, feat.double(), g.double(), lambda t_: t_)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    tol = FP32_TOL if dtype == torch.float32 else LOWP_TOL
    assert rel_err(out.float(), rout) <= tol
    assert rel_err(dfeat.float(), rdfeat) <= tol * (3 if dtype == torch.float32 else 1)  # head convs run in fp32 on the GPU side
    assert hg.keys() == rhg.keys() and len(hg) > 0
    for k in hg:
        assert rel_err(hg[k], rhg[k]) <= (1e-4 if dtype == torch.float32 else 2e-2), k
