"""GPU checks of the reference-facing surface around the kernels: the hook manager's two graph
semantics, the binary-mask path (downsample -> block with use_sigmoid_mask=False), autocast,
deepcopy (ModelEMA) and the flat gradient buffer."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import cbam_oracle as co
from oracle import mask_oracle as mo
from tests._golden import rel_err
from tests.test_host_logic import _ToyModel

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact_convs():
    """The toy model's own convs run through cuDNN; TF32 would amplify 1e-7 input differences to 1e-4."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _oracle_block(block, feat, mask, **kw):
    p = co.CbamParams.from_state_dict({k: v.detach().double().cpu() for k, v in block.state_dict().items()})
    out, _ = co.cbam_forward(feat.double().cpu(), None if mask is None else mask.double().cpu(), p,
                             use_sigmoid_mask=block.use_sigmoid_mask, feature_dtype=torch.float32, **kw)
    return out


def test_layer_output_hooks_feed_refined_features_forward():
    from mga_yolo_b200 import MGAHookManager

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = _ToyModel().to(dev).eval()
    mgr = MGAHookManager(net, target_layers=("1", "2", "3"), reduction_ratio=4)
    x = torch.randn(2, 3, 32, 32, device=dev)
    with torch.no_grad(), mgr:
        got = net(x)
        # manual walk: the refined output of layer i is what layer i+1 consumes
        y = net.model[0](x)
        feats = []
        for i in (1, 2, 3):
            handles = dict(net.model[i]._forward_hooks)
            net.model[i]._forward_hooks.clear()
            raw = net.model[i](y)
            net.model[i]._forward_hooks.update(handles)
            y = _oracle_block(mgr.blocks[str(i)], raw, None).float().to(dev)
            feats.append(y)
        net.model[4]._forward_hooks.clear()
        ref = net.model[4](feats)
    for a, b in zip(got, ref):
        assert rel_err(a, b) <= 1e-5


def test_detect_input_hook_matches_yaml_graph():
    from mga_yolo_b200 import MGAHookManager

    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    net = _ToyModel().to(dev).eval()
    mgr = MGAHookManager(net, target_layers=("1", "2", "3"), reduction_ratio=4, semantics="detect_input", use_sigmoid_mask=True)
    x = torch.randn(2, 3, 32, 32, device=dev)
    masks = [torch.randn(2, 1, s, s, device=dev) for s in (32, 16, 8)]
    with torch.no_grad():
        y = [net.model[0](x)]
        for i in (1, 2, 3):
            y.append(net.model[i](y[-1]))  # the PAN path keeps the RAW features
        refined = [_oracle_block(mgr.blocks[str(i)], y[i], masks[i - 1]).float().to(dev) for i in (1, 2, 3)]
        ref = net.model[4](refined)
        mgr.register()
        mgr.set_masks(masks)
        got = net(x)
        mgr.remove()
    for a, b in zip(got, ref):
        assert rel_err(a, b) <= 1e-5


@pytest.mark.parametrize("resize", ["nearest", "area", "maxpool"])
def test_binary_mask_path(resize):
    """image-size binary mask -> per-level downsample on the GPU -> block with use_sigmoid_mask=False"""
    from mga_yolo_b200 import MGAHookManager

    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    net = _ToyModel().to(dev).eval()
    mgr = MGAHookManager(net, target_layers=("1", "2", "3"), reduction_ratio=4, semantics="detect_input", mask_resize=resize)
    rng = np.random.default_rng(5)
    bm = (rng.random((2, 32, 32)) > 0.6).astype(np.uint8)
    x = torch.randn(2, 3, 32, 32, device=dev)
    with torch.no_grad():
        y = [net.model[0](x)]
        for i in (1, 2, 3):
            y.append(net.model[i](y[-1]))
        refined = []
        for i, stride in zip((1, 2, 3), (1, 2, 4)):
            if stride == 1:
                small = bm.astype(np.float32)
            elif resize == "area":
                small = np.stack([mo.area_u8(m, 32 // stride, 32 // stride) > 0 for m in bm]).astype(np.float32)
            else:
                small = np.stack([mo.downsample_mask(m, stride, resize, False) for m in bm]).astype(np.float32)
            refined.append(_oracle_block(mgr.blocks[str(i)], y[i], torch.from_numpy(small)[:, None]).float().to(dev))
        ref = net.model[4](refined)
        with mgr:
            mgr.set_masks(torch.from_numpy(bm).to(dev))
            got = net(x)
    for a, b in zip(got, ref):
        assert rel_err(a, b) <= 1e-5


def test_autocast_deepcopy_and_flat_grads():
    from mga_yolo_b200 import FlatGradReducer, MaskGuidedCBAM

    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    blk = MaskGuidedCBAM(64).to(dev)
    x = torch.randn(2, 64, 20, 20, device=dev)
    mk = torch.randn(2, 1, 20, 20, device=dev)
    with torch.autocast("cuda", dtype=torch.float16):  # the reference trainer's AMP mode: fp16 features, fp32 parameters
        out = blk([x.half(), mk.half()])
    assert out.dtype == torch.float16
    ref = _oracle_block(blk, x.half(), mk.half())
    assert rel_err(out.float().cpu(), ref) <= 1e-2
    ema = copy.deepcopy(blk)  # ModelEMA
    assert torch.equal(ema([x, mk]), blk([x, mk]))
    half = copy.deepcopy(blk).half()  # checkpoint.py:98 casts the whole model; parameters are read back as fp32
    assert half([x.half(), mk.half()]).dtype == torch.float16
    red = FlatGradReducer(blk.parameters())
    red.zero()
    xg = x.clone().requires_grad_(True)
    blk([xg, mk]).sum().backward()
    first = red.flat.clone()
    assert first.abs().sum() > 0 and blk.beta.grad.data_ptr() >= red.flat.data_ptr()
    blk([xg, mk]).sum().backward()  # autograd accumulates INTO the flat views
    assert torch.allclose(red.flat, 2 * first, rtol=1e-5, atol=1e-6)
    assert red.all_reduce() is None  # no process group: no-op


def test_gate_sampling_runs_on_gpu(monkeypatch):
    from mga_yolo_b200 import MaskGuidedCBAM

    monkeypatch.setenv("MGA_PROB_MODE", "1")
    monkeypatch.setenv("MGA_PROB_APPROACH", "gumbel")
    dev = torch.device("cuda:0")
    blk = MaskGuidedCBAM(32).to(dev).train()
    x = torch.randn(2, 32, 16, 16, device=dev, requires_grad=True)
    mk = torch.randn(2, 1, 16, 16, device=dev, requires_grad=True)
    out = blk([x, mk])
    out.sum().backward()
    assert torch.isfinite(out).all() and torch.isfinite(mk.grad).all() and mk.grad.abs().sum() > 0
