"""GPU mask downsample, bit-exact against the cv2 outputs the reference produced
(tests/golden/mask_downsample.npz) and against the numpy oracle at full image size."""
import numpy as np
import pytest
import torch

from oracle import mask_oracle as mo
from tests._golden import GOLDEN

pytestmark = pytest.mark.gpu


def test_bit_exact_vs_reference_goldens(monkeypatch):
    from mga_yolo_b200 import MaskUtils

    dev = torch.device("cuda:0")
    z = np.load(GOLDEN / "mask_downsample.npz")
    n = 0
    for si in range(len(z["sizes"])):
        src = torch.from_numpy(z[f"src{si}"]).to(dev)
        for s in z["strides"]:
            for method, bridge in (("nearest", "1"), ("area", "1"), ("area", "0"), ("maxpool", "1"),
                                   ("skeleton_bresenham", "1"), ("skeleton_bresenham", "0")):
                monkeypatch.setenv("MGA_MASK_METHOD", method)
                monkeypatch.setenv("MGA_MASK_BRIDGE", bridge)
                got = MaskUtils.downsample_mask(src, int(s))
                ref = torch.from_numpy(z[f"bin{si}_{s}_{method}_{bridge}"])
                assert got.dtype == torch.uint8 and torch.equal(got.cpu(), ref), (si, int(s), method, bridge)
                n += 1
            for method in ("avgpool", "nearest", "area"):
                got = MaskUtils.downsample_mask_prob(src, int(s), method)
                ref = torch.from_numpy(z[f"prob{si}_{s}_{method}"])
                assert got.dtype == torch.float32 and torch.equal(got.cpu(), ref), (si, int(s), method)
                n += 1
    assert n == len(z["sizes"]) * len(z["strides"]) * 9


@pytest.mark.parametrize("hw", [(640, 640), (1280, 1280), (608, 352)])
def test_batched_full_size_vs_oracle(hw, monkeypatch):
    from mga_yolo_b200 import MaskUtils

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(hw[0])
    B = 4
    H, W = hw
    src = (rng.random((B, H, W)) > 0.5).astype(np.uint8)
    src[:, : H // 2] &= (rng.random((B, H // 2, W)) > 0.7)
    d = torch.from_numpy(src).to(dev)
    for s in (8, 16, 32):
        for method, bridge in (("nearest", True), ("area", True), ("maxpool", False), ("skeleton_bresenham", True)):
            monkeypatch.setenv("MGA_MASK_METHOD", method)
            monkeypatch.setenv("MGA_MASK_BRIDGE", "1" if bridge else "0")
            got = MaskUtils.downsample_mask(d, s).cpu().numpy()
            for b in range(B):
                assert np.array_equal(got[b], mo.downsample_mask(src[b], s, method, bridge)), (hw, s, method, b)
        got = MaskUtils.downsample_mask_prob(d, s, "avgpool").cpu().numpy()
        for b in range(B):
            assert np.array_equal(got[b], mo.downsample_mask_prob(src[b], s, "avgpool"))
    multi = MaskUtils.masks_multi(d)
    assert [tuple(m.shape) for m in multi] == [(B, 1, -(-H // s), -(-W // s)) for s in (8, 16, 32)]


@pytest.mark.parametrize("B", [3, 9])  # 9: the two-stage form (whole-batch 8x8 block counts, then one CTA per image) of mga_masks_multi_ws
@pytest.mark.parametrize("hw", [(640, 640), (1280, 1280), (608, 352), (32, 64)])
def test_one_pass_multi_stride_is_bit_identical_to_per_stride(hw, B, monkeypatch):
    """torch.ops.mga.masks_multi (one read of the masks -> strides 8/16/32) against the per-stride op, which is pinned to the
    reference's cv2 outputs above, and against the numpy oracle; every method / bridge / output type of the dataset loop
    (mga_yolo/data/dataset.py:95-103)."""
    from mga_yolo_b200 import MaskUtils, _lib

    dev = torch.device("cuda:0")
    H, W = hw
    rng = np.random.default_rng(H + W)
    src = (rng.random((B, H, W)) > 0.55).astype(np.uint8)
    src[0, : H // 2] = 0
    src[1] &= (rng.random((H, W)) > 0.8)
    d = torch.from_numpy(src).to(dev)
    for method, bridge in (("nearest", "1"), ("area", "1"), ("area", "0"), ("maxpool", "1"), ("skeleton_bresenham", "1"), ("skeleton_bresenham", "0")):
        monkeypatch.setenv("MGA_MASK_METHOD", method)
        monkeypatch.setenv("MGA_MASK_BRIDGE", bridge)
        multi = MaskUtils.masks_multi(d)
        for m, s in zip(multi, (8, 16, 32)):
            assert m.dtype == torch.uint8 and m.shape == (B, 1, H // s, W // s)
            assert torch.equal(m[:, 0], MaskUtils.downsample_mask(d, s)), (hw, method, bridge, s)
            for b in range(B):
                assert np.array_equal(m[b, 0].cpu().numpy(), mo.downsample_mask(src[b], s, method, bridge == "1")), (hw, method, bridge, s, b)
    for method in ("avgpool", "nearest", "area"):
        monkeypatch.setenv("MGA_MASK_METHOD", method)
        multi = MaskUtils.masks_multi(d, prob=True)
        for m, s in zip(multi, (8, 16, 32)):
            assert m.dtype == torch.float32
            assert torch.equal(m[:, 0], MaskUtils.downsample_mask_prob(d, s, method)), (hw, method, s)
    # sizes the one-pass kernel does not take fall back to the per-stride path (same results by construction) ...
    odd = torch.from_numpy((rng.random((2, 100, 76)) > 0.5).astype(np.uint8)).to(dev)
    monkeypatch.setenv("MGA_MASK_METHOD", "maxpool")
    assert [tuple(m.shape) for m in MaskUtils.masks_multi(odd)] == [(2, 1, -(-100 // s), -(-76 // s)) for s in (8, 16, 32)]
    # ... and the C ABI says so instead of computing something else
    with pytest.raises(RuntimeError):
        torch.ops.mga.masks_multi(odd, _lib.DS_MAXPOOL, 0.0, False, False)
