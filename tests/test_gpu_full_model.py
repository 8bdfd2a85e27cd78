"""The drop-in, end to end, on the GPU: the reference's own `MGAModel('yolov8n_cbam.yaml')` -- built from baseline/_ref, the unmodified
reference package laid out by oracle/build_full_ref.py -- run forward + backward with its own classes, then again after
`mga_yolo_b200.install()` (MaskCBAM / MaskECA / MaskSPADE, MGAMaskHead swapped for the CUDA-library modules) with IDENTICAL weights and inputs.
What must agree: every output of `model(x)` ({"det": 3 maps, "seg": 3 logit maps}) and the gradients that flow through and out of the
swapped layers.  fp32 with TF32 off; differences are rounding-order only."""
import copy
import os
import sys
from pathlib import Path

import pytest
import torch

from tests._golden import rel_err

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"


def _flat(out):
    """{"det": [3 maps], "seg": {name: logits}} (mga_yolo/model/model.py:86) -> [(name, tensor)]"""
    res = []
    for k in ("det", "seg"):
        v = out[k]
        if isinstance(v, dict):
            res += [(f"{k}.{n}", v[n]) for n in sorted(v)]
        else:
            res += [(f"{k}[{i}]", t) for i, t in enumerate(v if isinstance(v, (list, tuple)) else [v])]
    return res


@pytest.mark.skipif(not (REF / "mga_yolo" / "__init__.py").exists(), reason="baseline/_ref not laid out (python oracle/build_full_ref.py)")
@pytest.mark.timeout(900)
@pytest.mark.parametrize("yaml_name", ["yolov8n_cbam.yaml", "yolov8n_eca.yaml", "yolov8n_spade.yaml"])
def test_reference_model_step_with_the_ops_swapped_in(yaml_name, monkeypatch):
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
    monkeypatch.delenv("MGA_PROB_MODE", raising=False)
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    monkeypatch.syspath_prepend(str(REF))
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    from mga_yolo.external.ultralytics.ultralytics import YOLO  # noqa: F401  (import order: SURVEY.md section 8c)
    from mga_yolo.model.model import MGAModel

    import mga_yolo_b200 as mb

    dev = torch.device("cuda:0")
    cfg = str(REF / "configs" / "models" / yaml_name)
    torch.manual_seed(0)
    ref = MGAModel(cfg, nc=1, verbose=False)
    assert type(ref.model[23]).__module__.startswith("mga_yolo.")
    with torch.no_grad():  # non-trivial gates: beta != 0, mask-head biases != 0
        for m in ref.model:
            if hasattr(m, "beta"):
                m.beta.fill_(0.3)
            if type(m).__name__ == "MGAMaskHead":
                m.head.bias.fill_(0.2)
    state = copy.deepcopy(ref.state_dict())
    x = torch.rand(2, 3, 160, 160, generator=torch.Generator().manual_seed(1)).to(dev)

    def step(model):
        model = model.to(dev).train()
        for p in model.parameters():
            p.grad = None
        xi = x.clone().requires_grad_(True)
        outs = _flat(model(xi))
        loss = sum(t.float().square().mean() for _, t in outs)
        loss.backward()
        grads = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.grad is not None}
        return [(n, t.detach().float().cpu()) for n, t in outs], xi.grad.float().cpu(), grads

    r_out, r_dx, r_g = step(ref)
    try:
        mb.install(strict=True)
        ours = MGAModel(cfg, nc=1, verbose=False)
        swapped = [type(m) for m in ours.model if isinstance(m, (mb.MaskCBAM, mb.MaskECA, mb.MaskSPADE, mb.MGAMaskHead))]
        assert len(swapped) == 6  # 3 mask heads + 3 attention blocks
        ours.load_state_dict(state, strict=True)  # same keys, same shapes: checkpoints interchange
        o_out, o_dx, o_g = step(ours)
    finally:
        mb.uninstall()
    assert [n for n, _ in o_out] == [n for n, _ in r_out] and len(r_out) == 6  # 3 detection maps + 3 mask-logit maps
    for (n, a), (_, b) in zip(o_out, r_out):
        assert a.shape == b.shape and rel_err(a, b) <= 2e-4, (n, rel_err(a, b))
    assert rel_err(o_dx, r_dx) <= 2e-3
    assert set(o_g) == set(r_g)
    worst = max((rel_err(o_g[n], r_g[n]), n) for n in r_g if r_g[n].abs().max() > 0)
    assert worst[0] <= 5e-3, worst
