"""Torch-facing operators of the components either side of the block:

    torch.ops.mga.eca_fwd / eca_bwd                   MaskECA                 (mga_yolo/nn/modules/masked_eca.py:139-193)
    torch.ops.mga.head_tail_fwd / head_tail_bwd       MGAMaskHead.head        (mga_yolo/nn/modules/segmentation.py:94,107-110)
    torch.ops.mga.gate_sample / gate_sample_bwd       ProbMaskGater, training (mga_yolo/nn/modules/probmaskgater.py:59-95)
    torch.ops.mga.collate_masks                       masks_multi collate     (mga_yolo/data/dataset.py:149-169)
    torch.ops.mga.spade_fwd / spade_bwd               MaskSPADE, feature side (mga_yolo/nn/modules/masked_spade.py:126-143)

CUDA dispatch key (the kernels of csrc/next_ops.cuh through the C ABI) and Meta key (shapes only): CPU tensors fail in the dispatcher.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .ops import _DT, _LIBDEF, _LIBIMPL, _stream

GATE_MODES = {"gumbel": 0, "hard_st": 1, "bernoulli_detach": 2}

_LIBDEF.define("eca_fwd(Tensor x, Tensor? mask, Tensor w1d, Tensor beta, int flags, float tiny_thr, float eps) -> (Tensor, Tensor)")
_LIBDEF.define("eca_bwd(Tensor grad_out, Tensor x, Tensor? mask, Tensor w1d, Tensor ctx, int flags, float tiny_thr, float eps, bool need_mask_grad)"
               " -> (Tensor, Tensor?, Tensor, Tensor)")
_LIBDEF.define("head_tail_fwd(Tensor feat, Tensor weight, Tensor bias) -> Tensor")
_LIBDEF.define("head_tail_bwd(Tensor grad_logits, Tensor feat, Tensor weight) -> (Tensor, Tensor, Tensor)")
_LIBDEF.define("gate_sample(Tensor p, Tensor? noise, int mode, float tau, float p_min, float threshold, int seed, int offset) -> (Tensor, Tensor)")
_LIBDEF.define("gate_sample_bwd(Tensor grad_out, Tensor p, Tensor soft, float tau, float p_min) -> Tensor")
_LIBDEF.define("collate_masks(Tensor[] maps) -> Tensor")
_LIBDEF.define("spade_fwd(Tensor x, Tensor? gamma, Tensor? beta, float eps) -> (Tensor, Tensor)")
_LIBDEF.define("spade_bwd(Tensor grad_out, Tensor x, Tensor? gamma, Tensor stats, bool need_gamma_grad) -> (Tensor, Tensor?)")


@lru_cache(maxsize=256)
def _eca_desc(B, Cc, H, W, k, dt, mdt, flags, tiny, eps):
    d = _lib.Desc(B, Cc, H, W, k, 1, dt, mdt, flags, tiny, eps)
    cb, sb = C.c_size_t(0), C.c_size_t(0)
    _lib.check(_lib.load().mga_eca_workspace(C.byref(d), C.byref(cb), C.byref(sb)), "mga_eca_workspace")
    return d, int(cb.value), int(sb.value)


def _eca_prep(x, mask, w1d, flags, tiny, eps):
    if x.dim() != 4 or x.dtype not in _DT:
        raise RuntimeError(f"feature map must be (B,C,H,W) float32 / bfloat16 / float16, got {tuple(x.shape)} {x.dtype}")
    B, Cc, H, W = x.shape
    mdt = _lib.F32
    if mask is not None:
        if mask.dtype not in _DT:
            raise RuntimeError(f"unsupported mask dtype {mask.dtype}")
        if mask.numel() != B * H * W or tuple(mask.shape[-2:]) != (H, W):
            raise RuntimeError(f"mask {tuple(mask.shape)} does not match feature map {tuple(x.shape)}")  # masked_eca.py:149 expand fails too
        mdt = _DT[mask.dtype]
        flags |= _lib.HAS_MASK
    else:
        flags &= ~_lib.HAS_MASK
    return _eca_desc(B, Cc, H, W, int(w1d.numel()), _DT[x.dtype], mdt, flags, float(tiny), float(eps))


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _eca_fwd_cuda(x, mask, w1d, beta, flags, tiny_thr, eps):
    lib = _lib.load()
    x = x.contiguous()
    mask = None if mask is None else mask.contiguous()
    d, ctx_bytes, scratch_bytes = _eca_prep(x, mask, w1d, flags, tiny_thr, eps)
    w, b = _f32c(w1d.reshape(-1)), _f32c(beta)
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        ctx = torch.empty(ctx_bytes, dtype=torch.uint8, device=x.device)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=x.device)
        rc = lib.mga_eca_forward(C.byref(d), x.data_ptr(), None if mask is None else mask.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(),
                                 ctx.data_ptr(), scratch.data_ptr(), _stream(x))
    _lib.check(rc, "mga_eca_forward")
    return out, ctx


def _eca_bwd_cuda(grad_out, x, mask, w1d, ctx, flags, tiny_thr, eps, need_mask_grad):
    lib = _lib.load()
    x = x.contiguous()
    grad_out = grad_out.contiguous().to(x.dtype)
    mask = None if mask is None else mask.contiguous()
    d, _, scratch_bytes = _eca_prep(x, mask, w1d, flags, tiny_thr, eps)
    w = _f32c(w1d.reshape(-1))
    with torch.cuda.device(x.device):
        dx = torch.empty_like(x)
        dmask = torch.empty_like(mask) if (mask is not None and need_mask_grad) else None
        dw = torch.empty(w.numel(), dtype=torch.float32, device=x.device)
        dbeta = torch.empty((), dtype=torch.float32, device=x.device)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=x.device)
        rc = lib.mga_eca_backward(C.byref(d), x.data_ptr(), None if mask is None else mask.data_ptr(), grad_out.data_ptr(), w.data_ptr(),
                                  ctx.data_ptr(), dx.data_ptr(), None if dmask is None else dmask.data_ptr(), dw.data_ptr(), dbeta.data_ptr(),
                                  scratch.data_ptr(), _stream(x))
    _lib.check(rc, "mga_eca_backward")
    return dx, dmask, dw, dbeta


def _head_tail_fwd_cuda(feat, weight, bias):
    lib = _lib.load()
    if feat.dim() != 4 or feat.dtype not in _DT:
        raise RuntimeError(f"hidden feature must be (B,C,H,W) float32 / bfloat16 / float16, got {tuple(feat.shape)} {feat.dtype}")
    feat = feat.contiguous()
    B, Cc, H, W = feat.shape
    if tuple(weight.shape) != (1, Cc, 3, 3):
        raise RuntimeError(f"head weight must be (1,{Cc},3,3), got {tuple(weight.shape)}")
    w, b = _f32c(weight), _f32c(bias)
    with torch.cuda.device(feat.device):
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=feat.device)
        rc = lib.mga_head_tail_forward(feat.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, Cc, H, W, _DT[feat.dtype], _stream(feat))
    _lib.check(rc, "mga_head_tail_forward")
    return out


def _head_tail_bwd_cuda(grad_logits, feat, weight):
    lib = _lib.load()
    feat = feat.contiguous()
    B, Cc, H, W = feat.shape
    g, w = _f32c(grad_logits), _f32c(weight)
    with torch.cuda.device(feat.device):
        dfeat = torch.empty_like(feat)
        dw = torch.empty((1, Cc, 3, 3), dtype=torch.float32, device=feat.device)
        db = torch.empty((1,), dtype=torch.float32, device=feat.device)
        rc = lib.mga_head_tail_backward(feat.data_ptr(), w.data_ptr(), g.data_ptr(), dfeat.data_ptr(), dw.data_ptr(), db.data_ptr(), B, Cc, H, W,
                                        _DT[feat.dtype], _stream(feat))
    _lib.check(rc, "mga_head_tail_backward")
    return dfeat, dw, db


def _gate_sample_cuda(p, noise, mode, tau, p_min, threshold, seed, offset):
    lib = _lib.load()
    p = _f32c(p)
    n = p.numel()
    if noise is not None:
        noise = _f32c(noise)
        if noise.numel() != 2 * n:
            raise RuntimeError("noise must hold two uniforms per element: shape (2, *p.shape)")
    with torch.cuda.device(p.device):
        out = torch.empty_like(p)
        soft = torch.empty_like(p)
        rc = lib.mga_gate_sample_forward(p.data_ptr(), None if noise is None else noise.data_ptr(), out.data_ptr(), soft.data_ptr(), None, n, int(mode),
                                         float(tau), float(p_min), float(threshold), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), _stream(p))
    _lib.check(rc, "mga_gate_sample_forward")
    return out, soft


def _gate_sample_bwd_cuda(grad_out, p, soft, tau, p_min):
    lib = _lib.load()
    g, p = _f32c(grad_out), _f32c(p)
    with torch.cuda.device(p.device):
        dp = torch.empty_like(p)
        rc = lib.mga_gate_sample_backward(g.data_ptr(), p.data_ptr(), soft.data_ptr(), dp.data_ptr(), p.numel(), float(tau), float(p_min), _stream(p))
    _lib.check(rc, "mga_gate_sample_backward")
    return dp


def _collate_cuda(maps: List[torch.Tensor]):
    lib = _lib.load()
    if not maps:
        raise RuntimeError("collate_masks needs at least one map")
    dev = maps[0].device
    dt = maps[0].dtype
    if dt not in (torch.uint8, torch.float32) or any(m.dtype != dt or m.device != dev for m in maps):
        raise RuntimeError("collate_masks takes uint8 or float32 maps of one dtype on one device")
    maps = [m.reshape(m.shape[-2], m.shape[-1]).contiguous() for m in maps]
    H, W = max(m.shape[0] for m in maps), max(m.shape[1] for m in maps)
    rec = np.zeros(len(maps), dtype=np.dtype([("src", np.uint64), ("h", np.int32), ("w", np.int32)]))
    for i, m in enumerate(maps):
        rec[i] = (m.data_ptr(), m.shape[0], m.shape[1])
    with torch.cuda.device(dev):
        items = torch.from_numpy(rec.view(np.uint8)).to(dev, non_blocking=False)
        out = torch.empty((len(maps), 1, H, W), dtype=torch.float32, device=dev)
        rc = lib.mga_collate_masks(items.data_ptr(), out.data_ptr(), len(maps), H, W, _lib.F32 if dt == torch.float32 else _lib.U8, _stream(out))
    _lib.check(rc, "mga_collate_masks")
    return out


def _spade_prep(x, gamma, beta=None):
    if x.dim() != 4 or x.dtype not in _DT:
        raise RuntimeError(f"feature map must be (B,C,H,W) float32 / bfloat16 / float16, got {tuple(x.shape)} {x.dtype}")
    for name, t in (("gamma", gamma), ("beta", beta)):
        if t is None:
            continue
        if tuple(t.shape) != tuple(x.shape):
            raise RuntimeError(f"{name} {tuple(t.shape)} does not match the feature map {tuple(x.shape)}")
        if t.dtype not in (x.dtype, torch.float32):
            raise RuntimeError(f"{name} must have the feature dtype or float32, got {t.dtype}")
    if beta is not None and gamma is not None and beta.dtype != gamma.dtype:
        raise RuntimeError("gamma and beta must share a dtype")
    return _DT[x.dtype], (_DT[gamma.dtype] if gamma is not None else _DT[x.dtype])


def _spade_fwd_cuda(x, gamma, beta, eps):
    lib = _lib.load()
    if (gamma is None) != (beta is None):
        raise RuntimeError("gamma and beta come together")
    dt, mdt = _spade_prep(x, gamma, beta)
    x = x.contiguous()
    gamma = None if gamma is None else gamma.contiguous()
    beta = None if beta is None else beta.contiguous()
    B, Cc, H, W = x.shape
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        stats = torch.empty((B, Cc, 2), dtype=torch.float32, device=x.device)
        rc = lib.mga_spade_forward(x.data_ptr(), None if gamma is None else gamma.data_ptr(), None if beta is None else beta.data_ptr(), out.data_ptr(),
                                   stats.data_ptr(), B, Cc, H, W, float(eps), dt, mdt, _stream(x))
    _lib.check(rc, "mga_spade_forward")
    return out, stats


def _spade_bwd_cuda(grad_out, x, gamma, stats, need_gamma_grad):
    lib = _lib.load()
    dt, mdt = _spade_prep(x, gamma)
    x = x.contiguous()
    grad_out = grad_out.contiguous().to(x.dtype)
    gamma = None if gamma is None else gamma.contiguous()
    B, Cc, H, W = x.shape
    with torch.cuda.device(x.device):
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(gamma) if (gamma is not None and need_gamma_grad) else None
        rc = lib.mga_spade_backward(x.data_ptr(), grad_out.data_ptr(), None if gamma is None else gamma.data_ptr(), stats.data_ptr(), dx.data_ptr(),
                                    None if dgamma is None else dgamma.data_ptr(), B, Cc, H, W, dt, mdt, _stream(x))
    _lib.check(rc, "mga_spade_backward")
    return dx, dgamma


# ---- Meta kernels (shapes only)
def _eca_fwd_meta(x, mask, w1d, beta, flags, tiny_thr, eps):
    _, ctx_bytes, _ = _eca_prep(x, mask, w1d, flags, tiny_thr, eps)
    return torch.empty_like(x, memory_format=torch.contiguous_format), x.new_empty(ctx_bytes, dtype=torch.uint8)


def _eca_bwd_meta(grad_out, x, mask, w1d, ctx, flags, tiny_thr, eps, need_mask_grad):
    dmask = torch.empty_like(mask, memory_format=torch.contiguous_format) if (mask is not None and need_mask_grad) else None
    return (torch.empty_like(x, memory_format=torch.contiguous_format), dmask, x.new_empty(w1d.numel(), dtype=torch.float32),
            x.new_empty((), dtype=torch.float32))


def _head_tail_fwd_meta(feat, weight, bias):
    B, _, H, W = feat.shape
    return feat.new_empty((B, 1, H, W), dtype=torch.float32)


def _head_tail_bwd_meta(grad_logits, feat, weight):
    return torch.empty_like(feat, memory_format=torch.contiguous_format), feat.new_empty(weight.shape, dtype=torch.float32), feat.new_empty((1,), dtype=torch.float32)


def _gate_sample_meta(p, noise, mode, tau, p_min, threshold, seed, offset):
    return p.new_empty(p.shape, dtype=torch.float32), p.new_empty(p.shape, dtype=torch.float32)


def _gate_sample_bwd_meta(grad_out, p, soft, tau, p_min):
    return p.new_empty(p.shape, dtype=torch.float32)


def _collate_meta(maps):
    H, W = max(m.shape[-2] for m in maps), max(m.shape[-1] for m in maps)
    return maps[0].new_empty((len(maps), 1, H, W), dtype=torch.float32)


def _spade_fwd_meta(x, gamma, beta, eps):
    _spade_prep(x, gamma, beta)
    return torch.empty_like(x, memory_format=torch.contiguous_format), x.new_empty((x.shape[0], x.shape[1], 2), dtype=torch.float32)


def _spade_bwd_meta(grad_out, x, gamma, stats, need_gamma_grad):
    dgamma = torch.empty_like(gamma, memory_format=torch.contiguous_format) if (gamma is not None and need_gamma_grad) else None
    return torch.empty_like(x, memory_format=torch.contiguous_format), dgamma


for _name, _cuda, _meta in (("eca_fwd", _eca_fwd_cuda, _eca_fwd_meta), ("eca_bwd", _eca_bwd_cuda, _eca_bwd_meta),
                            ("head_tail_fwd", _head_tail_fwd_cuda, _head_tail_fwd_meta), ("head_tail_bwd", _head_tail_bwd_cuda, _head_tail_bwd_meta),
                            ("gate_sample", _gate_sample_cuda, _gate_sample_meta), ("gate_sample_bwd", _gate_sample_bwd_cuda, _gate_sample_bwd_meta),
                            ("collate_masks", _collate_cuda, _collate_meta), ("spade_fwd", _spade_fwd_cuda, _spade_fwd_meta),
                            ("spade_bwd", _spade_bwd_cuda, _spade_bwd_meta)):
    _LIBIMPL.impl(_name, _cuda, "CUDA")
    _LIBIMPL.impl(_name, _meta, "Meta")


def _need_cuda(t, what):
    if t.device.type not in ("cuda", "meta"):
        raise RuntimeError(f"mga_yolo_b200: {what} runs on CUDA tensors only (no CPU fallback)")


class _EcaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask, w1d, beta, flags, tiny_thr, eps):
        _need_cuda(x, "MaskECA")
        out, saved = torch.ops.mga.eca_fwd(x, mask, w1d, beta, flags, tiny_thr, eps)
        ctx.save_for_backward(x, mask, w1d, saved)
        ctx.cfg = (flags, tiny_thr, eps)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, mask, w1d, saved = ctx.saved_tensors
        flags, tiny_thr, eps = ctx.cfg
        need_mask = mask is not None and ctx.needs_input_grad[1]
        dx, dmask, dw, dbeta = torch.ops.mga.eca_bwd(grad_out, x, mask, w1d, saved, flags, tiny_thr, eps, need_mask)
        return dx, dmask, dw.view(w1d.shape), dbeta, None, None, None


def mask_eca(x, mask, w1d, beta, *, flags: int, tiny_mask_thr: float = 1e-4, eps: float = 1e-6):
    """out = MaskECA([x, mask]) with the given conv1d weight (1,1,k) and beta (autograd-aware)."""
    return _EcaFn.apply(x, mask, w1d, beta, int(flags), float(tiny_mask_thr), float(eps))


class _SpadeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        _need_cuda(x, "MaskSPADE")
        out, stats = torch.ops.mga.spade_fwd(x, gamma, beta, eps)
        ctx.save_for_backward(x, gamma, stats)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, gamma, stats = ctx.saved_tensors
        need_gamma = gamma is not None and ctx.needs_input_grad[1]
        dx, dgamma = torch.ops.mga.spade_bwd(grad_out, x, gamma, stats, need_gamma)
        dbeta = None
        if gamma is not None and ctx.needs_input_grad[2]:
            dbeta = grad_out if grad_out.dtype == gamma.dtype else grad_out.to(gamma.dtype)  # d beta IS grad_out: nothing to compute
        return dx, dgamma, dbeta, None


def spade_modulate(x: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float = 1e-6) -> torch.Tensor:
    """gamma * InstanceNorm(x) + beta (plain InstanceNorm when gamma is None) by the CUDA kernels of csrc/spade_ops.cu (autograd-aware)."""
    return _SpadeFn.apply(x, gamma, beta, float(eps))


class _HeadTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, weight, bias):
        _need_cuda(feat, "the MGAMaskHead tail")
        ctx.save_for_backward(feat, weight)
        return torch.ops.mga.head_tail_fwd(feat, weight, bias)

    @staticmethod
    def backward(ctx, grad_logits):
        feat, weight = ctx.saved_tensors
        dfeat, dw, db = torch.ops.mga.head_tail_bwd(grad_logits, feat, weight)
        return dfeat, dw.to(weight.dtype), db


def head_tail(feat, weight, bias):
    """Mask logits (B,1,H,W) fp32 = Conv2d(hidden, 1, 3, padding=1)(feat) by the CUDA tail kernel (autograd-aware)."""
    return _HeadTailFn.apply(feat, weight, bias)


class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, noise, mode, tau, p_min, threshold, seed, offset):
        _need_cuda(p, "ProbMaskGater sampling")
        out, soft = torch.ops.mga.gate_sample(p, noise, mode, tau, p_min, threshold, seed, offset)
        ctx.save_for_backward(p, soft)
        ctx.cfg = (mode, tau, p_min)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        p, soft = ctx.saved_tensors
        mode, tau, p_min = ctx.cfg
        if mode == GATE_MODES["bernoulli_detach"]:
            return (None,) * 8
        return (torch.ops.mga.gate_sample_bwd(grad_out, p, soft, tau, p_min).view(p.shape).to(p.dtype),) + (None,) * 7


def gate_sample(p: torch.Tensor, mode: str, *, tau: float = 1.0, p_min: float = 0.0, threshold: float = 0.5, seed: int = 0, offset: int = 0,
                noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Train-mode ProbMaskGater on a (B,1,H,W) map; noise contract in include/mga_cbam.h (Philox4x32-10 keyed by seed, offset)."""
    return _GateFn.apply(p, noise, GATE_MODES[mode], float(tau), float(p_min), float(threshold), int(seed), int(offset))


# ---------------------------------------------------------------- sam_cam_fusion = concat: fused forward on the tensor cores
_LIBDEF.define("cbam_concat_fwd(Tensor x, Tensor s, Tensor a, Tensor w, Tensor bias, Tensor beta, bool pyramid_multiply) -> Tensor")


def concat_fused_supported(x: torch.Tensor) -> bool:
    """Shapes / types mga_cbam_concat_forward takes (include/mga_cbam.h)."""
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.bfloat16, torch.float16) and x.shape[1] % 128 == 0
            and (x.shape[2] * x.shape[3]) % 8 == 0)


def _concat_fwd_cuda(x, s, a, w, bias, beta, pyramid_multiply):
    lib = _lib.load()
    x = x.contiguous()
    B, Cc, H, W = x.shape
    if tuple(w.shape[:2]) != (Cc, 2 * Cc):
        raise RuntimeError(f"fuse_sam_cam weight must be ({Cc},{2 * Cc},1,1), got {tuple(w.shape)}")
    s, a, w, bias, beta = _f32c(s), _f32c(a), _f32c(w.reshape(Cc, 2 * Cc)), _f32c(bias), _f32c(beta)
    d = _lib.Desc(B, Cc, H, W, 1, 1, _DT[x.dtype], _lib.F32, _lib.PYRAMID_MULTIPLY if pyramid_multiply else 0, 0.0, 0.0)
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        ws = torch.empty((B + 1) * Cc * Cc, dtype=x.dtype, device=x.device)
        rc = lib.mga_cbam_concat_forward(C.byref(d), x.data_ptr(), s.data_ptr(), a.data_ptr(), w.data_ptr(), bias.data_ptr(), beta.data_ptr(),
                                         out.data_ptr(), ws.data_ptr(), _stream(x))
    _lib.check(rc, "mga_cbam_concat_forward")
    return out


def _concat_fwd_meta(x, s, a, w, bias, beta, pyramid_multiply):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


_LIBIMPL.impl("cbam_concat_fwd", _concat_fwd_cuda, "CUDA")
_LIBIMPL.impl("cbam_concat_fwd", _concat_fwd_meta, "Meta")


class _ConcatFn(torch.autograd.Function):
    """out = k0 x + k1 (Wa (x s) + Wb (x a) + bias).  Forward: ONE tcgen05 kernel (+ the weight-folding pre-kernel).  Backward: closed form
    = three library GEMMs (U|V = [Wa^T ; Wb^T] g; g X^T; (g a) X^T) + ONE elementwise / reduction kernel (mga_cbam_concat_backward_elem);
    the 2C-channel concat tensor is not formed in either direction."""

    @staticmethod
    def forward(ctx, x, s, a, w, bias, beta, pyramid_multiply):
        out = torch.ops.mga.cbam_concat_fwd(x, s, a, w, bias, beta, pyramid_multiply)
        ctx.save_for_backward(x, s, a, w, bias, beta)
        ctx.pm = pyramid_multiply
        return out

    @staticmethod
    def backward(ctx, gout):
        x, s, a, w, bias, beta = ctx.saved_tensors
        dx, ds, da, dw, dbias, dbeta = _concat_backward(x, s, a, w, bias, beta, gout, ctx.pm)
        return dx, ds, da, dw, dbias, dbeta, None


def _concat_backward(x, s, a, w, bias, beta, gout, pyramid_multiply):
    """Closed-form backward of cbam_concat_fwd: (dx, ds, da, dw, dbias, dbeta).
    Default: ONE tcgen05 kernel (mga_cbam_concat_backward_dx: U = Wa^T g and V = Wb^T g as its two accumulators, the elementwise /
    reduction part of the closed form in its epilogue), two per-sample library GEMMs (g X^T, (g a) X^T) and ONE batch-reduce kernel for
    the weight gradient; narrow levels (C <= 256) take U|V from a library GEMM + mga_cbam_concat_backward_elem (see the dispatch note)."""
    import os

    lib = _lib.load()
    B, Cc, H, W = x.shape
    S = H * W
    dt = x.dtype
    x = x.contiguous()
    g = gout.contiguous().to(dt)
    w2 = _f32c(w.reshape(Cc, 2 * Cc))
    sf, af, bf, btf = _f32c(s), _f32c(a), _f32c(bias), _f32c(beta)
    d = _lib.Desc(B, Cc, H, W, 1, 1, _DT[dt], _lib.F32, _lib.PYRAMID_MULTIPLY if pyramid_multiply else 0, 0.0, 0.0)
    with torch.cuda.device(x.device):
        dx = torch.empty_like(x)
        ga = torch.empty_like(x)
        # Measured on B200 (cfg4, B = 128 bf16, tools/concat_bwd_prof.py): the tcgen05 backward beats library GEMM + elementwise kernel for
        # C > 256 (512x40x40: 474 vs 522 us, 512x20x20: 134 vs 164 us); at C <= 256 the epilogue of the resident-weight kernel (8 warps, one
        # accumulator row per lane) is the bound (256x80x80: 912 vs 766 us), so that width keeps the library form.  MGA_CONCAT_BWD = tc | library forces one.
        force = os.getenv("MGA_CONCAT_BWD", "")
        if force == "library" or (force != "tc" and Cc <= 256):
            wcat_t = torch.cat([w2[:, :Cc].t(), w2[:, Cc:].t()], dim=0).to(dt).contiguous()  # (2C, C): [Wa^T ; Wb^T]
            uv = torch.matmul(wcat_t, g.reshape(B, Cc, S))                                     # (B, 2C, S), one library GEMM
            nT = (S // 8 + 31) // 32
            ds_part = torch.empty((B, nT, Cc), dtype=torch.float32, device=x.device)
            db_part = torch.empty((B, nT, Cc), dtype=torch.float32, device=x.device)
            da = torch.empty((B, S), dtype=torch.float32, device=x.device)
            dal_part = torch.empty((B, nT), dtype=torch.float32, device=x.device)
            rc = lib.mga_cbam_concat_backward_elem(C.byref(d), x.data_ptr(), g.data_ptr(), uv.data_ptr(), sf.data_ptr(), af.data_ptr(), bf.data_ptr(),
                                                   btf.data_ptr(), dx.data_ptr(), ga.data_ptr(), ds_part.data_ptr(), db_part.data_ptr(),
                                                   da.data_ptr(), dal_part.data_ptr(), _stream(x))
            _lib.check(rc, "mga_cbam_concat_backward_elem")
            del uv
        else:
            nT = (S + 127) // 128
            ew = int(os.getenv("MGA_CC_EPI_WARPS", "8"))  # epilogue warps of the library build (tuning builds may carry 16)
            ds_part = torch.empty((B, (ew // 4) * nT, Cc), dtype=torch.float32, device=x.device)
            db_part = torch.empty((B, (ew // 4) * nT, Cc), dtype=torch.float32, device=x.device)
            da_part = torch.empty((B, Cc // 32, S), dtype=torch.float32, device=x.device)
            dal_part = torch.empty((B, nT, (Cc // 128) * ew), dtype=torch.float32, device=x.device)
            ws = torch.empty(2 * Cc * Cc, dtype=dt, device=x.device)
            rc = lib.mga_cbam_concat_backward_dx(C.byref(d), x.data_ptr(), g.data_ptr(), sf.data_ptr(), af.data_ptr(), w2.data_ptr(), bf.data_ptr(),
                                                 btf.data_ptr(), dx.data_ptr(), ga.data_ptr(), ds_part.data_ptr(), db_part.data_ptr(),
                                                 da_part.data_ptr(), dal_part.data_ptr(), ws.data_ptr(), _stream(x))
            _lib.check(rc, "mga_cbam_concat_backward_dx")
            da = da_part.sum(dim=1)
        # weight gradient: two per-sample GEMMs with fp32 results, then ONE kernel sums them over the batch (the channel gate folded in)
        xt = x.reshape(B, Cc, S).transpose(1, 2)
        Ga = torch.bmm(g.reshape(B, Cc, S), xt, out_dtype=torch.float32)                       # (B,C,C): g X^T
        Gb = torch.bmm(ga.reshape(B, Cc, S), xt, out_dtype=torch.float32)                      # (g * a) X^T
        dw32 = torch.empty((Cc, 2 * Cc), dtype=torch.float32, device=x.device)
        rc = lib.mga_cbam_concat_wgrad_reduce(C.byref(d), Ga.data_ptr(), Gb.data_ptr(), _lib.F32, sf.data_ptr(), btf.data_ptr(), dw32.data_ptr(),
                                              _stream(x))
        _lib.check(rc, "mga_cbam_concat_wgrad_reduce")
    dw = dw32.reshape(w.shape).to(w.dtype)
    ds = ds_part.sum(dim=1).reshape(s.shape).to(s.dtype)
    dbias = db_part.sum(dim=(0, 1)).to(bias.dtype)
    dbeta = (torch.sigmoid(btf) * dal_part.double().sum().float()).reshape(beta.shape).to(beta.dtype)
    return dx, ds, da.reshape(a.shape).to(a.dtype), dw, dbias, dbeta


class _ConcatBlockFn(torch.autograd.Function):
    """The whole `sam_cam_fusion = concat` block -- gates op + fused concat -- as ONE autograd node: the feature gradient through the two
    1x1 convolutions and the one through the gates are summed inside the kernel that writes grad_x (mga_cbam_gates_backward_acc), not by
    an extra pass over two N-sized tensors."""

    @staticmethod
    def forward(ctx, x, mask, w1, b1, w2, b2, wsam, wf, bf, beta, flags, tiny_thr, eps, pyramid_multiply):
        s, a, saved = torch.ops.mga.cbam_gates_fwd(x, mask, w1, b1, w2, b2, wsam, flags, tiny_thr, eps)
        out = torch.ops.mga.cbam_concat_fwd(x, s, a, wf, bf, beta, pyramid_multiply)
        ctx.save_for_backward(x, mask, w1, b1, w2, b2, wsam, wf, bf, beta, s, a, saved)
        ctx.cfg = (flags, tiny_thr, eps, pyramid_multiply)
        ctx.param_shapes = tuple(t.shape for t in (w1, b1, w2, b2, wsam))
        return out

    @staticmethod
    def backward(ctx, gout):
        x, mask, w1, b1, w2, b2, wsam, wf, bf, beta, s, a, saved = ctx.saved_tensors
        flags, tiny_thr, eps, pm = ctx.cfg
        dx_c, ds, da, dwf, dbf, dbeta = _concat_backward(x, s, a, wf, bf, beta, gout, pm)
        need_mask = mask is not None and ctx.needs_input_grad[1]
        dx, dmask, flat = torch.ops.mga.cbam_gates_bwd(ds, da, x, mask, w1, b1, w2, b2, wsam, saved, flags, tiny_thr, eps, need_mask, dx_c)
        grads, o = [], 0
        for shp in ctx.param_shapes:
            n = 1
            for v in shp:
                n *= v
            grads.append(flat[o:o + n].view(shp))
            o += n
        return (dx, dmask, *grads, dwf, dbf, dbeta, None, None, None, None)


def concat_block(x, mask, w1, b1, w2, b2, wsam, wf, bf, beta, *, flags: int, tiny_mask_thr: float, eps: float, pyramid_multiply: bool):
    """out = k0 x + k1 (Wa (x s) + Wb (x a') + bias) with the gates (s, a') of the block computed from (x, mask): gates op + tcgen05 concat
    kernel forward, one autograd node (see _ConcatBlockFn)."""
    return _ConcatBlockFn.apply(x, mask, w1, b1, w2, b2, wsam, wf, bf, beta, int(flags), float(tiny_mask_thr), float(eps), bool(pyramid_multiply))


def concat_fused(x, s, a, w, bias, beta, pyramid_multiply: bool):
    """Fused `sam_cam_fusion=concat` forward (tcgen05) with a library-GEMM closed-form backward; see include/mga_cbam.h."""
    return _ConcatFn.apply(x, s, a, w, bias, beta, bool(pyramid_multiply))
