"""Torch-facing operators of the mask-guided CBAM path.

`torch.ops.mga.cbam_fwd / cbam_bwd / mask_downsample` are registered for the CUDA
dispatch key (the kernels) and the Meta key (shape propagation only), so CPU tensors fail
inside the dispatcher ("no CPU fallback").
Each op is a thin shim: it allocates outputs with the caching allocator and hands raw
device pointers + the current stream to the C ABI (include/mga_cbam.h).

Replaces, on the reference side, the body of MaskCBAM.forward
(mga_yolo/nn/modules/masked_cbam.py:154-171) and its autograd graph.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache
from typing import Optional, Tuple

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}

_LIBDEF = torch.library.Library("mga", "DEF")
_LIBDEF.define(
    "cbam_fwd(Tensor x, Tensor? mask, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor wsam, Tensor beta, "
    "int flags, float tiny_thr, float eps) -> (Tensor, Tensor)"
)
_LIBDEF.define(
    "cbam_bwd(Tensor grad_out, Tensor x, Tensor? mask, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor wsam, Tensor beta, "
    "Tensor ctx, int flags, float tiny_thr, float eps, bool need_mask_grad) -> (Tensor, Tensor?, Tensor)"
)
_LIBDEF.define(
    "cbam_gates_fwd(Tensor x, Tensor? mask, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor wsam, int flags, float tiny_thr, "
    "float eps) -> (Tensor, Tensor, Tensor)"
)
_LIBDEF.define(
    "cbam_gates_bwd(Tensor grad_s, Tensor grad_a, Tensor x, Tensor? mask, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor wsam, "
    "Tensor ctx, int flags, float tiny_thr, float eps, bool need_mask_grad, Tensor? grad_x_acc=None) -> (Tensor, Tensor?, Tensor)"
)
_LIBDEF.define("mask_downsample(Tensor src, int stride, int method, float thresh, bool close3x3, bool out_float) -> Tensor")
_LIBDEF.define("masks_multi(Tensor src, int method, float thresh, bool close3x3, bool out_float) -> (Tensor, Tensor, Tensor)")


@lru_cache(maxsize=256)
def _desc(B, Cc, H, W, hidden, k, dt, mdt, flags, tiny, eps):
    d = _lib.Desc(B, Cc, H, W, hidden, k, dt, mdt, flags, tiny, eps)
    lib = _lib.load()
    cb, sb = C.c_size_t(0), C.c_size_t(0)
    _lib.check(lib.mga_cbam_workspace(C.byref(d), C.byref(cb), C.byref(sb)), "mga_cbam_workspace")
    return d, int(cb.value), int(sb.value)


def _prep(x, mask, w1, wsam, flags, tiny, eps):
    if x.dim() != 4:
        raise RuntimeError(f"feature map must be (B,C,H,W), got {tuple(x.shape)}")
    if x.dtype not in _DT:
        raise RuntimeError(f"unsupported feature dtype {x.dtype} (float32 / bfloat16 / float16)")
    B, Cc, H, W = x.shape
    hidden, k = w1.shape[0], wsam.shape[-1]
    if tuple(w1.shape) != (hidden, Cc) or tuple(wsam.shape) != (1, 3, k, k):
        raise RuntimeError(f"parameter shapes do not match C={Cc}: w1 {tuple(w1.shape)}, wsam {tuple(wsam.shape)}")
    mdt = _lib.F32
    if mask is not None:
        if mask.dtype not in _DT:
            raise RuntimeError(f"unsupported mask dtype {mask.dtype}")
        if mask.numel() != B * H * W or tuple(mask.shape[-2:]) != (H, W):
            # the reference fails in mask.expand(b,c,h,w) (masked_cbam.py:96) for any other size
            raise RuntimeError(f"mask {tuple(mask.shape)} does not match feature map {tuple(x.shape)}")
        mdt = _DT[mask.dtype]
        flags |= _lib.HAS_MASK
    else:
        flags &= ~_lib.HAS_MASK
    return _desc(B, Cc, H, W, hidden, k, _DT[x.dtype], mdt, flags, float(tiny), float(eps))


def _params(w1, b1, w2, b2, wsam, beta):
    ts = [t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous() for t in (w1, b1, w2, b2, wsam, beta)]
    return ts, _lib.Params(*(t.data_ptr() for t in ts))


def _stream(x) -> int:
    return torch.cuda.current_stream(x.device).cuda_stream


def _cbam_fwd_cuda(x, mask, w1, b1, w2, b2, wsam, beta, flags, tiny_thr, eps):
    lib = _lib.load()
    x = x.contiguous()
    mask = None if mask is None else mask.contiguous()
    d, ctx_bytes, scratch_bytes = _prep(x, mask, w1, wsam, flags, tiny_thr, eps)
    keep, prm = _params(w1, b1, w2, b2, wsam, beta)
    with torch.cuda.device(x.device):
        out = torch.empty_like(x)
        ctx = torch.empty(ctx_bytes, dtype=torch.uint8, device=x.device)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=x.device)
        rc = lib.mga_cbam_forward(C.byref(d), x.data_ptr(), None if mask is None else mask.data_ptr(), C.byref(prm),
                                  out.data_ptr(), ctx.data_ptr(), scratch.data_ptr(), _stream(x))
    _lib.check(rc, "mga_cbam_forward")
    return out, ctx


def _cbam_bwd_cuda(grad_out, x, mask, w1, b1, w2, b2, wsam, beta, ctx, flags, tiny_thr, eps, need_mask_grad):
    lib = _lib.load()
    x = x.contiguous()
    grad_out = grad_out.contiguous()
    if grad_out.dtype != x.dtype:
        grad_out = grad_out.to(x.dtype)
    mask = None if mask is None else mask.contiguous()
    d, _, scratch_bytes = _prep(x, mask, w1, wsam, flags, tiny_thr, eps)
    keep, prm = _params(w1, b1, w2, b2, wsam, beta)
    hidden, Cc, k = w1.shape[0], x.shape[1], wsam.shape[-1]
    n = [hidden * Cc, hidden, Cc * hidden, Cc, 3 * k * k, 1]
    with torch.cuda.device(x.device):
        dx = torch.empty_like(x)
        dmask = torch.empty_like(mask) if (mask is not None and need_mask_grad) else None
        flat = torch.empty(sum(n), dtype=torch.float32, device=x.device)
        offs = [0]
        for v in n:
            offs.append(offs[-1] + v)
        gp = _lib.Grads(*(flat.data_ptr() + 4 * o for o in offs[:-1]))
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=x.device)
        rc = lib.mga_cbam_backward(C.byref(d), x.data_ptr(), None if mask is None else mask.data_ptr(), grad_out.data_ptr(),
                                   C.byref(prm), ctx.data_ptr(), dx.data_ptr(), None if dmask is None else dmask.data_ptr(),
                                   C.byref(gp), scratch.data_ptr(), _stream(x))
    _lib.check(rc, "mga_cbam_backward")
    return dx, dmask, flat


def _mask_downsample_cuda(src, stride, method, thresh, close3x3, out_float):
    lib = _lib.load()
    if src.dtype != torch.uint8:
        raise RuntimeError("mask_downsample expects a uint8 {0,1} mask (mask_utils.py:26-27,81-82 binarise first)")
    squeeze = src.dim() == 2
    s3 = (src[None] if squeeze else src).contiguous()
    if s3.dim() != 3:
        raise RuntimeError(f"mask must be (H,W) or (B,H,W), got {tuple(src.shape)}")
    B, H, W = s3.shape
    if stride <= 1:
        out = s3.float() if out_float else s3.clone()
        return out[0] if squeeze else out
    nh, nw = -(-H // stride), -(-W // stride)
    with torch.cuda.device(src.device):
        out = torch.empty((B, nh, nw), dtype=torch.float32 if out_float else torch.uint8, device=src.device)
        tmp = torch.empty(2 * B * nh * nw, dtype=torch.uint8, device=src.device) if close3x3 else None
        rc = lib.mga_mask_downsample(s3.data_ptr(), out.data_ptr(), None if tmp is None else tmp.data_ptr(), B, H, W, stride,
                                     method, float(thresh), int(close3x3), _lib.F32 if out_float else _lib.U8, _stream(src))
    _lib.check(rc, "mga_mask_downsample")
    return out[0] if squeeze else out


def masks_multi_supported(H: int, W: int) -> bool:
    """Shapes the one-pass kernel takes (include/mga_cbam.h: mga_masks_multi)."""
    return H % 32 == 0 and W % 32 == 0 and (H // 8) * (W // 8) <= 32768


def _masks_multi_cuda(src, method, thresh, close3x3, out_float):
    """(B,H,W) uint8 {0,1} -> three maps at strides 8 / 16 / 32 from ONE read of the masks (dataset.py:95-103)."""
    lib = _lib.load()
    if src.dtype != torch.uint8 or src.dim() != 3:
        raise RuntimeError("masks_multi expects a (B,H,W) uint8 {0,1} batch of masks")
    s3 = src.contiguous()
    B, H, W = s3.shape
    if not masks_multi_supported(H, W):
        raise RuntimeError(f"masks_multi needs H, W divisible by 32 (got {H}x{W}); use mask_downsample per stride")
    dt = torch.float32 if out_float else torch.uint8
    with torch.cuda.device(src.device):
        outs = [torch.empty((B, H // s, W // s), dtype=dt, device=src.device) for s in (8, 16, 32)]
        # two stages from 8 images on: one thread per 8x8 block reads the masks (grid over the whole batch), then one CTA per image
        tmp = torch.empty(2 * B * (H // 8) * (W // 8), dtype=torch.uint8, device=src.device) if B >= 8 else None
        rc = lib.mga_masks_multi_ws(s3.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), None if tmp is None else tmp.data_ptr(),
                                    B, H, W, method, float(thresh), int(close3x3), _lib.F32 if out_float else _lib.U8, _stream(src))
    _lib.check(rc, "mga_masks_multi")
    return tuple(outs)


def _cbam_gates_fwd_cuda(x, mask, w1, b1, w2, b2, wsam, flags, tiny_thr, eps):
    lib = _lib.load()
    x = x.contiguous()
    mask = None if mask is None else mask.contiguous()
    d, ctx_bytes, scratch_bytes = _prep(x, mask, w1, wsam, flags, tiny_thr, eps)
    with torch.cuda.device(x.device):
        beta = torch.zeros((), dtype=torch.float32, device=x.device)  # unused by the gates
        keep, prm = _params(w1, b1, w2, b2, wsam, beta)
        ctx = torch.empty(ctx_bytes, dtype=torch.uint8, device=x.device)
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=x.device)
        rc = lib.mga_cbam_gates_forward(C.byref(d), x.data_ptr(), None if mask is None else mask.data_ptr(), C.byref(prm),
                                        ctx.data_ptr(), scratch.data_ptr(), _stream(x))
    _lib.check(rc, "mga_cbam_gates_forward")
    B, Cc, H, W = x.shape
    views = []
    for which, shape in ((0, (B, Cc)), (1, (B, 1, H, W))):
        ptr, cnt = C.c_void_p(0), C.c_size_t(0)
        _lib.check(lib.mga_cbam_ctx_view(C.byref(d), ctx.data_ptr(), which, C.byref(ptr), C.byref(cnt)), "mga_cbam_ctx_view")
        off = ptr.value - ctx.data_ptr()
        views.append(ctx[off:off + 4 * cnt.value].view(torch.float32).view(shape).clone())
    return views[0], views[1], ctx


def _cbam_gates_bwd_cuda(grad_s, grad_a, x, mask, w1, b1, w2, b2, wsam, ctx, flags, tiny_thr, eps, need_mask_grad, grad_x_acc=None):
    lib = _lib.load()
    x = x.contiguous()
    if grad_x_acc is not None:  # upstream feature gradient: summed into dx by the kernel that writes it (no extra pass)
        if grad_x_acc.shape != x.shape:
            raise RuntimeError(f"grad_x_acc {tuple(grad_x_acc.shape)} does not match x {tuple(x.shape)}")
        grad_x_acc = grad_x_acc.to(x.dtype).contiguous()
    mask = None if mask is None else mask.contiguous()
    grad_s = grad_s.float().contiguous()
    grad_a = grad_a.float().contiguous()
    d, _, scratch_bytes = _prep(x, mask, w1, wsam, flags, tiny_thr, eps)
    hidden, Cc, k = w1.shape[0], x.shape[1], wsam.shape[-1]
    n = [hidden * Cc, hidden, Cc * hidden, Cc, 3 * k * k, 1]
    with torch.cuda.device(x.device):
        beta = torch.zeros((), dtype=torch.float32, device=x.device)
        keep, prm = _params(w1, b1, w2, b2, wsam, beta)
        dx = torch.empty_like(x)
        dmask = torch.empty_like(mask) if (mask is not None and need_mask_grad) else None
        flat = torch.empty(sum(n), dtype=torch.float32, device=x.device)
        offs = [0]
        for v in n:
            offs.append(offs[-1] + v)
        gp = _lib.Grads(*(flat.data_ptr() + 4 * o for o in offs[:-1]))
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=x.device)
        rc = lib.mga_cbam_gates_backward_acc(C.byref(d), x.data_ptr(), None if mask is None else mask.data_ptr(), grad_s.data_ptr(),
                                             grad_a.data_ptr(), None if grad_x_acc is None else grad_x_acc.data_ptr(), C.byref(prm),
                                             ctx.data_ptr(), dx.data_ptr(), None if dmask is None else dmask.data_ptr(), C.byref(gp),
                                             scratch.data_ptr(), _stream(x))
    _lib.check(rc, "mga_cbam_gates_backward_acc")
    return dx, dmask, flat


# ---- Meta ("fake") kernels: shape / dtype propagation only, so meta tensors, FakeTensorMode and graph capture see the ops.
def _ctx_bytes_of(x, mask, w1, wsam, flags, tiny_thr, eps):
    d, ctx_bytes, _ = _prep(x, mask, w1, wsam, flags, tiny_thr, eps)
    return ctx_bytes


def _grad_numel(x, w1, wsam):
    hidden, Cc, k = w1.shape[0], x.shape[1], wsam.shape[-1]
    return hidden * Cc + hidden + Cc * hidden + Cc + 3 * k * k + 1


def _cbam_fwd_meta(x, mask, w1, b1, w2, b2, wsam, beta, flags, tiny_thr, eps):
    return torch.empty_like(x, memory_format=torch.contiguous_format), x.new_empty(_ctx_bytes_of(x, mask, w1, wsam, flags, tiny_thr, eps), dtype=torch.uint8)


def _cbam_bwd_meta(grad_out, x, mask, w1, b1, w2, b2, wsam, beta, ctx, flags, tiny_thr, eps, need_mask_grad):
    dmask = torch.empty_like(mask, memory_format=torch.contiguous_format) if (mask is not None and need_mask_grad) else None
    return torch.empty_like(x, memory_format=torch.contiguous_format), dmask, x.new_empty(_grad_numel(x, w1, wsam), dtype=torch.float32)


def _cbam_gates_fwd_meta(x, mask, w1, b1, w2, b2, wsam, flags, tiny_thr, eps):
    B, Cc, H, W = x.shape
    return (x.new_empty((B, Cc), dtype=torch.float32), x.new_empty((B, 1, H, W), dtype=torch.float32),
            x.new_empty(_ctx_bytes_of(x, mask, w1, wsam, flags, tiny_thr, eps), dtype=torch.uint8))


def _cbam_gates_bwd_meta(grad_s, grad_a, x, mask, w1, b1, w2, b2, wsam, ctx, flags, tiny_thr, eps, need_mask_grad, grad_x_acc=None):
    dmask = torch.empty_like(mask, memory_format=torch.contiguous_format) if (mask is not None and need_mask_grad) else None
    return torch.empty_like(x, memory_format=torch.contiguous_format), dmask, x.new_empty(_grad_numel(x, w1, wsam), dtype=torch.float32)


def _mask_downsample_meta(src, stride, method, thresh, close3x3, out_float):
    dt = torch.float32 if out_float else torch.uint8
    if stride <= 1:
        return src.new_empty(src.shape, dtype=dt)
    H, W = src.shape[-2:]
    return src.new_empty((*src.shape[:-2], -(-H // stride), -(-W // stride)), dtype=dt)


def _masks_multi_meta(src, method, thresh, close3x3, out_float):
    B, H, W = src.shape
    dt = torch.float32 if out_float else torch.uint8
    return tuple(src.new_empty((B, H // s, W // s), dtype=dt) for s in (8, 16, 32))


_LIBIMPL = torch.library.Library("mga", "IMPL")
for _name, _fn in (("cbam_fwd", _cbam_fwd_meta), ("cbam_bwd", _cbam_bwd_meta), ("cbam_gates_fwd", _cbam_gates_fwd_meta),
                   ("cbam_gates_bwd", _cbam_gates_bwd_meta), ("mask_downsample", _mask_downsample_meta), ("masks_multi", _masks_multi_meta)):
    _LIBIMPL.impl(_name, _fn, "Meta")
_LIBIMPL.impl("cbam_gates_fwd", _cbam_gates_fwd_cuda, "CUDA")
_LIBIMPL.impl("cbam_gates_bwd", _cbam_gates_bwd_cuda, "CUDA")
_LIBIMPL.impl("cbam_fwd", _cbam_fwd_cuda, "CUDA")
_LIBIMPL.impl("cbam_bwd", _cbam_bwd_cuda, "CUDA")
_LIBIMPL.impl("mask_downsample", _mask_downsample_cuda, "CUDA")
_LIBIMPL.impl("masks_multi", _masks_multi_cuda, "CUDA")


class _CbamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask, w1, b1, w2, b2, wsam, beta, flags, tiny_thr, eps):
        if x.device.type not in ("cuda", "meta"):
            raise RuntimeError("mga_yolo_b200: the mask-guided CBAM path runs on CUDA tensors only (no CPU fallback)")
        out, saved = torch.ops.mga.cbam_fwd(x, mask, w1, b1, w2, b2, wsam, beta, flags, tiny_thr, eps)
        ctx.save_for_backward(x, mask, w1, b1, w2, b2, wsam, beta, saved)
        ctx.cfg = (flags, tiny_thr, eps)
        ctx.param_shapes = tuple(t.shape for t in (w1, b1, w2, b2, wsam, beta))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, mask, w1, b1, w2, b2, wsam, beta, saved = ctx.saved_tensors
        flags, tiny_thr, eps = ctx.cfg
        need_mask = mask is not None and ctx.needs_input_grad[1]
        dx, dmask, flat = torch.ops.mga.cbam_bwd(grad_out, x, mask, w1, b1, w2, b2, wsam, beta, saved, flags, tiny_thr, eps, need_mask)
        grads, o = [], 0
        for shp in ctx.param_shapes:
            n = 1
            for v in shp:
                n *= v
            grads.append(flat[o:o + n].view(shp))
            o += n
        return (dx, dmask, *grads, None, None, None)


class _CbamGatesFn(torch.autograd.Function):
    """(s, a) = the channel gate (B,C) and the spatial gate computed from x (B,1,H,W), both fp32."""

    @staticmethod
    def forward(ctx, x, mask, w1, b1, w2, b2, wsam, flags, tiny_thr, eps):
        if x.device.type not in ("cuda", "meta"):
            raise RuntimeError("mga_yolo_b200: the mask-guided CBAM path runs on CUDA tensors only (no CPU fallback)")
        s, a, saved = torch.ops.mga.cbam_gates_fwd(x, mask, w1, b1, w2, b2, wsam, flags, tiny_thr, eps)
        ctx.save_for_backward(x, mask, w1, b1, w2, b2, wsam, saved)
        ctx.cfg = (flags, tiny_thr, eps)
        ctx.param_shapes = tuple(t.shape for t in (w1, b1, w2, b2, wsam))
        return s, a

    @staticmethod
    def backward(ctx, grad_s, grad_a):
        x, mask, w1, b1, w2, b2, wsam, saved = ctx.saved_tensors
        flags, tiny_thr, eps = ctx.cfg
        need_mask = mask is not None and ctx.needs_input_grad[1]
        dx, dmask, flat = torch.ops.mga.cbam_gates_bwd(grad_s, grad_a, x, mask, w1, b1, w2, b2, wsam, saved, flags, tiny_thr, eps, need_mask)
        grads, o = [], 0
        for shp in ctx.param_shapes:
            n = 1
            for v in shp:
                n *= v
            grads.append(flat[o:o + n].view(shp))
            o += n
        return (dx, dmask, *grads, None, None, None)


def cbam_gates(x, mask, w1, b1, w2, b2, wsam, *, flags: int, tiny_mask_thr: float = 1e-4, eps: float = 1e-6):
    """Differentiable gates of the block: s = sigmoid(MLP(masked avg) + MLP(masked max)) (B,C) and a' = sigmoid(conv7x7([max_c x,
    mean_c x, m])) (B,1,H,W).  Used by the `concat` fusion modes, whose 1x1 convolutions are ordinary library GEMMs."""
    return _CbamGatesFn.apply(x, mask, w1, b1, w2, b2, wsam, int(flags), float(tiny_mask_thr), float(eps))


def mask_guided_cbam(x: torch.Tensor, mask: Optional[torch.Tensor], w1, b1, w2, b2, wsam, beta, *, flags: int,
                     tiny_mask_thr: float = 1e-4, eps: float = 1e-6) -> torch.Tensor:
    """Functional entry: out = MaskCBAM([x, mask]) with the given parameters (autograd-aware)."""
    needs_grad = torch.is_grad_enabled() and any(
        isinstance(t, torch.Tensor) and t.requires_grad for t in (x, mask, w1, b1, w2, b2, wsam, beta))
    if not needs_grad:  # inference (predictor / validator): nothing is saved for a backward that will never run
        flags = int(flags) | _lib.NO_SAVE
    return _CbamFn.apply(x, mask, w1, b1, w2, b2, wsam, beta, int(flags), float(tiny_mask_thr), float(eps))


def plan(x_shape: Tuple[int, int, int, int], dtype=torch.float32, *, r: int = 16, k: int = 7, flags: int = 0, backward: bool = False) -> dict:
    """Which launch path (and cluster geometry) the library takes for a feature map of this shape: host-only, no GPU needed.
    {'path': 'cluster' | 'per_phase', 'cluster_size', 'rows_per_cta', 'threads', 'smem_bytes', 'launches'}."""
    lib = _lib.load()
    B, Cc, H, W = x_shape
    d = _lib.Desc(B, Cc, H, W, max(1, Cc // r), k, _DT[dtype], _lib.F32, int(flags) | _lib.HAS_MASK | _lib.SIGMOID_MASK, 1e-4, 1e-6)
    info = _lib.PlanInfo()
    _lib.check(lib.mga_cbam_plan(C.byref(d), 1 if backward else 0, C.byref(info)), "mga_cbam_plan")
    out = {n: int(getattr(info, n)) for n, _ in _lib.PlanInfo._fields_}
    out["path"] = "cluster" if info.path == 1 else "per_phase"
    return out


def ctx_view(x_shape: Tuple[int, int, int, int], dtype, hidden: int, k: int, flags: int, ctx: torch.Tensor, which: int,
             tiny=1e-4, eps=1e-6, mask_dtype=torch.float32) -> torch.Tensor:
    """Copy a small saved quantity out of a forward context (tests / logging): 0 s(B,C), 1 a(B,HW)."""
    lib = _lib.load()
    B, Cc, H, W = x_shape
    d, _, _ = _desc(B, Cc, H, W, hidden, k, _DT[dtype], _DT[mask_dtype], flags, float(tiny), float(eps))
    ptr, cnt = C.c_void_p(0), C.c_size_t(0)
    _lib.check(lib.mga_cbam_ctx_view(C.byref(d), ctx.data_ptr(), which, C.byref(ptr), C.byref(cnt)), "mga_cbam_ctx_view")
    off = (ptr.value - ctx.data_ptr())
    return ctx[off:off + 4 * cnt.value].view(torch.float32).clone()
