"""CPU oracle for the binary-mask downsample path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Restates, in integer / float32 numpy, what the reference obtains from OpenCV
(`opencv-python>=4.6.0`, reference pyproject.toml:28; 4.13.0 in this image; the
library is a third-party dependency that is not vendored in /root/reference):

  MaskUtils.downsample_mask       mga_yolo/utils/mask_utils.py:64-141
      nearest  -> cv2.INTER_NEAREST                       (:90-91)
      area     -> cv2.INTER_AREA, > thresh, 3x3 close     (:93-99)
      maxpool  -> zero pad, block max                     (:101-110)
      default  -> "skeleton_bresenham", non-strict: block max + 3x3 close
                  (mga_yolo/utils/mask_downsample.py:135-145,159-165)
  MaskUtils.downsample_mask_prob  mga_yolo/utils/mask_utils.py:14-48
      avgpool / nearest / area (uint8 in -> uint8 {0,1} out)

Parity status: PINNED -- tests/golden/mask_*.npz were produced by calling the
reference functions (and therefore cv2) in oracle/gen_golden.py; this restatement
is checked bit-for-bit against them in tests/test_oracle_golden.py.

Published algorithms restated here (OpenCV imgproc/resize.cpp):
  INTER_NEAREST : src index = min(floor(dst * (ssize / dsize)), ssize - 1), double math.
  INTER_AREA, integer scale (both axes): int block sum * float32(1/area), then
      round-half-even saturate to uint8.
  INTER_AREA, fractional scale: separable float32 table of (dst, src, weight)
      entries (computeResizeAreaTab), horizontal pass into a float row, vertical
      accumulation, round-half-even saturate.
  MORPH_CLOSE 3x3 : dilate then erode, out-of-image neighbours ignored.
"""
from __future__ import annotations

import math

import numpy as np

METHODS = ("nearest", "area", "maxpool", "skeleton_bresenham")
PROB_METHODS = ("avgpool", "nearest", "area")


def out_size(h: int, w: int, stride: int):
    return math.ceil(h / stride), math.ceil(w / stride)


def _binarise(mask: np.ndarray) -> np.ndarray:
    # mask_utils.py:26-27, 81-82: anything that is not uint8 becomes (mask > 0)
    if mask.dtype != np.uint8:
        return (mask > 0).astype(np.uint8)
    return mask


def nearest(mask: np.ndarray, nh: int, nw: int) -> np.ndarray:
    h, w = mask.shape
    iy = np.minimum(np.floor(np.arange(nh, dtype=np.float64) * (h / nh)).astype(np.int64), h - 1)
    ix = np.minimum(np.floor(np.arange(nw, dtype=np.float64) * (w / nw)).astype(np.int64), w - 1)
    return mask[iy][:, ix]


def _area_table(ssize: int, dsize: int):
    """(dst, src, float32 weight) triples in emission order (computeResizeAreaTab)."""
    scale = ssize / dsize
    tab = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1 = math.ceil(f1)
        s2 = math.floor(f2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            tab.append((d, s1 - 1, np.float32((s1 - f1) / cell)))
        for sx in range(s1, s2):
            tab.append((d, sx, np.float32(1.0 / cell)))
        if f2 - s2 > 1e-3:
            tab.append((d, s2, np.float32(min(min(f2 - s2, 1.0), cell) / cell)))
    return tab


def _round_half_even_u8(v: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def area_u8(mask: np.ndarray, nh: int, nw: int) -> np.ndarray:
    """cv2.resize(uint8, INTER_AREA) for a down-scale."""
    h, w = mask.shape
    sy, sx = h / nh, w / nw
    if abs(sy - round(sy)) < 2.2e-16 and abs(sx - round(sx)) < 2.2e-16:
        ky, kx = int(round(sy)), int(round(sx))
        blocks = mask[: nh * ky, : nw * kx].reshape(nh, ky, nw, kx).astype(np.int32).sum(axis=(1, 3))
        scale = np.float32(1.0) / np.float32(ky * kx)
        return _round_half_even_u8(blocks.astype(np.float32) * scale)
    xtab = _area_table(w, nw)
    ytab = _area_table(h, nh)
    src = mask.astype(np.float32)
    out = np.zeros((nh, nw), dtype=np.uint8)
    acc = np.zeros(nw, dtype=np.float32)
    prev = ytab[0][0]
    first = True
    for (dy, syi, beta) in ytab:
        buf = np.zeros(nw, dtype=np.float32)
        for (dx, sxi, alpha) in xtab:
            buf[dx] = np.float32(buf[dx] + src[syi, sxi] * alpha)
        if dy != prev:
            out[prev] = _round_half_even_u8(acc)
            prev = dy
            first = True
        if first:
            acc = (beta * buf).astype(np.float32)
            first = False
        else:
            acc = (acc + beta * buf).astype(np.float32)
    out[prev] = _round_half_even_u8(acc)
    return out


def block_reduce(mask: np.ndarray, stride: int, op: str) -> np.ndarray:
    h, w = mask.shape
    ph, pw = (-h) % stride, (-w) % stride
    if ph or pw:
        mask = np.pad(mask, ((0, ph), (0, pw)))
    hh, ww = mask.shape
    view = mask.reshape(hh // stride, stride, ww // stride, stride)
    if op == "max":
        return view.max(axis=(1, 3)).astype(np.uint8)
    return view.astype(np.float32).mean(axis=(1, 3)).astype(np.float32)


def close3x3(img: np.ndarray) -> np.ndarray:
    def nbr(a, fill, fn):
        h, w = a.shape
        p = np.full((h + 2, w + 2), fill, dtype=a.dtype)
        p[1:-1, 1:-1] = a
        r = p[1:-1, 1:-1].copy()
        for dy in range(3):
            for dx in range(3):
                r = fn(r, p[dy : dy + h, dx : dx + w])
        return r

    d = nbr(img, 0, np.maximum)  # dilate: border contributes nothing
    return nbr(d, 255, np.minimum)  # erode: border contributes nothing


def downsample_mask(mask: np.ndarray, stride: int, method: str = "skeleton_bresenham", bridge: bool = True, thresh: float = 0.0) -> np.ndarray:
    mask = _binarise(mask)
    if stride <= 1:
        return mask
    nh, nw = out_size(*mask.shape, stride)
    if method == "nearest":
        return nearest(mask, nh, nw)
    if method == "area":
        out = (area_u8(mask, nh, nw) > thresh).astype(np.uint8)
        return close3x3(out) if bridge else out
    if method == "maxpool":
        return block_reduce(mask, stride, "max")
    if method == "skeleton_bresenham":  # non-strict default: occupancy + optional close
        out = block_reduce((mask > 0).astype(np.uint8), stride, "max")
        return close3x3(out) if bridge else out
    raise ValueError(f"method {method!r} is outside the hot path")


def downsample_mask_prob(mask: np.ndarray, stride: int, method: str = "area") -> np.ndarray:
    if stride <= 1:
        return mask.astype(np.float32)
    mask = _binarise(mask)
    nh, nw = out_size(*mask.shape, stride)
    if method == "avgpool":
        return block_reduce(mask, stride, "mean")
    if method == "nearest":
        return nearest(mask, nh, nw).astype(np.float32)
    return np.clip(area_u8(mask, nh, nw).astype(np.float32), 0.0, 1.0)
