"""`MaskGuidedCBAM` -- the nn.Module face of the B200 mask-guided CBAM path.

Drop-in for the reference block `MaskCBAM` (mga_yolo/nn/modules/masked_cbam.py:10-174):
same constructor arguments, same state_dict keys and shapes
(`beta`, `cam_mlp.0.{weight,bias}`, `cam_mlp.2.{weight,bias}`, `sam_conv.weight`), same
`forward(x)` polymorphism (`Tensor` or `[feature, mask]`), same `.alpha` property that
MGATrainer logs (mga_yolo/model/trainer.py:274-321).  The north-star keyword names
(`reduction_ratio`, `sam_cam_fusion`, `mga_pyramid_fusion`) are accepted too.

The forward/backward math does not live here: it is one call into the CUDA library
through `ops.mask_guided_cbam`.  CPU tensors raise -- there is no compute fallback.

Two non-CUDA cases are answered WITHOUT computing anything, because graph builders ask for
shapes before the model is on a GPU:
  * meta tensors -> `torch.ops.mga.*` Meta kernels (shape/dtype propagation only);
  * inside `shape_probe()` -- which `install()` wraps around the reference's
    `DetectionModel.__init__` (ultralytics/nn/tasks.py:418-426: a CPU forward of zeros whose only
    use is `s / x.shape[-2]` per output level) -- a CPU feature map gets an all-zeros tensor of
    its own shape and dtype.  Outside that context a CPU tensor is an error.
"""
from __future__ import annotations

import contextlib
import os
import threading
from typing import Optional, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, next_ops, ops

_GATE_MODES = ("deterministic", "gumbel", "hard_st", "bernoulli_detach")
# sam_cam_fusion: how the channel gate s and the spatial gate a combine
#   multiply -> x*s*a with a computed from x*s            (== reference, masked_cbam.py:166-168)
#   add      -> x*s + x*a' with a' computed from x        (build-side definition, parity unpinned)
# mga_pyramid_fusion: how the refined feature re-enters the pyramid
#   add      -> x + alpha*(R - x)                         (== reference, masked_cbam.py:169-171)
#   multiply -> alpha*R                                   (build-side definition, parity unpinned)
#   concat   -> Conv1x1([x*s ; x*a']) 2C->C with an extra learnable layer `fuse_sam_cam`      (build-side, parity unpinned)
# mga_pyramid_fusion = concat -> Conv1x1([x ; R]) 2C->C with an extra learnable layer `fuse_pyramid`   (build-side, parity unpinned)
# The concat modes run as: CUDA gates op (s, a') + elementwise torch ops + a cuDNN/cuBLAS 1x1 convolution (a plain library GEMM).
SAM_CAM_FUSIONS = ("multiply", "add", "concat")
PYRAMID_FUSIONS = ("add", "multiply", "concat")
_BETA_ALPHA_ONE = 0.5413248546129181  # softplus(.) == 1: lets the fused op return the bare refined feature R = x*s*a

_probe = threading.local()


@contextlib.contextmanager
def shape_probe():
    """While active (per thread), a CPU forward of the block answers with zeros of the input's shape: the stride probe of
    the reference's graph builder (ultralytics/nn/tasks.py:418-426) only reads output shapes.  Nothing is computed."""
    _probe.depth = getattr(_probe, "depth", 0) + 1
    try:
        yield
    finally:
        _probe.depth -= 1


def in_shape_probe() -> bool:
    return getattr(_probe, "depth", 0) > 0


class MaskGate(nn.Module):
    """Mask pre-gate used when the MGA_PROB_MODE environment switch is on
    (probmaskgater.py:27-98).  The deterministic branch (eval mode, or mode
    'deterministic') is only a clamp to [0,1] and is fused into the CUDA kernels; the
    sampling branches are ONE CUDA kernel (Philox noise + logistic-gumbel sigmoid / straight-through / Bernoulli) with a closed-form backward."""

    def __init__(self, mode: str = "gumbel", tau: float = 1.0, p_min: float = 0.0, threshold: float = 0.5):
        super().__init__()
        if tau <= 0:
            raise ValueError("tau must be > 0")
        self.mode, self.tau, self.p_min, self.threshold = mode, float(tau), float(p_min), float(threshold)
        self.seed: Optional[int] = None  # Philox key of this gate's noise stream (set on first use)
        self.calls = 0                   # stream offset: one per sampled forward

    def is_deterministic(self) -> bool:
        return (not self.training) or self.mode == "deterministic" or self.mode not in _GATE_MODES

    def sample(self, mask: torch.Tensor) -> torch.Tensor:
        """Train-mode gate (probmaskgater.py:73-95) by the CUDA sampling kernel.  Noise contract (include/mga_cbam.h): Philox4x32-10
        keyed by `self.seed` with the call counter as the stream offset -- the reference draws from torch's global generator with
        seed=None (masked_cbam.py:74-78), so its samples are not reproducible either; parity is on the distribution and, with the
        reference's own uniforms fed through `noise`, on the values (tests/test_gpu_next.py)."""
        from . import next_ops

        p = mask if mask.dim() == 4 else mask.unsqueeze(1)
        if self.seed is None:  # fresh stream per module instance, drawn once from torch's generator (so torch.manual_seed governs it)
            self.seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        out = next_ops.gate_sample(p.float(), self.mode, tau=self.tau, p_min=self.p_min, threshold=self.threshold, seed=self.seed, offset=self.calls)
        self.calls += 1
        return out


class MaskGuidedCBAM(nn.Module):
    def __init__(
        self,
        channels: int,
        r: Optional[int] = None,
        spatial_k: int = 7,
        use_sigmoid_mask: bool = True,
        tiny_mask_thr: float = 1e-4,
        eps: float = 1e-6,
        *,
        reduction_ratio: Optional[int] = None,
        sam_cam_fusion: str = "multiply",
        mga_pyramid_fusion: str = "add",
    ) -> None:
        super().__init__()
        if r is not None and reduction_ratio is not None and r != reduction_ratio:
            raise ValueError("pass either r or reduction_ratio, not two different values")
        r = reduction_ratio if r is None and reduction_ratio is not None else (16 if r is None else r)
        assert r > 0 and channels > 0
        if sam_cam_fusion not in SAM_CAM_FUSIONS:
            raise ValueError(f"sam_cam_fusion must be one of {SAM_CAM_FUSIONS}, got {sam_cam_fusion!r}")
        if mga_pyramid_fusion not in PYRAMID_FUSIONS:
            raise ValueError(f"mga_pyramid_fusion must be one of {PYRAMID_FUSIONS}, got {mga_pyramid_fusion!r}")
        self.C = channels
        self.r = r
        self.k = spatial_k if spatial_k % 2 == 1 else spatial_k + 1
        self.use_sigmoid_mask = use_sigmoid_mask
        self.tiny_thr = tiny_mask_thr
        self.eps = eps
        self.sam_cam_fusion = sam_cam_fusion
        self.mga_pyramid_fusion = mga_pyramid_fusion

        # parameter containers only -- created in the reference's order (masked_cbam.py:53-64) so that
        # the same torch seed gives the same initial weights; they are never called as layers.
        hidden = max(1, channels // r)
        self.cam_mlp = nn.Sequential(nn.Linear(channels, hidden, bias=True), nn.ReLU(inplace=True), nn.Linear(hidden, channels, bias=True))
        self.sam_conv = nn.Conv2d(3, 1, kernel_size=self.k, padding=self.k // 2, bias=False)
        self.beta = nn.Parameter(torch.zeros((), dtype=torch.float32))
        # extra layers exist only in the concat modes, so the default state_dict stays the reference's
        if sam_cam_fusion == "concat":
            self.fuse_sam_cam = nn.Conv2d(2 * channels, channels, kernel_size=1, bias=True)
        if mga_pyramid_fusion == "concat":
            self.fuse_pyramid = nn.Conv2d(2 * channels, channels, kernel_size=1, bias=True)

        # any non-empty MGA_PROB_MODE string switches the gate on, like os.getenv(..., False) does (masked_cbam.py:67)
        if os.getenv("MGA_PROB_MODE", False):
            approach = os.getenv("MGA_PROB_APPROACH", "gumbel")
            if approach not in _GATE_MODES:
                raise ValueError(f"MGA_PROB_APPROACH must be one of {set(_GATE_MODES)}, got {approach}")
            self.gater = MaskGate(mode=approach, tau=1.0, p_min=0.0, threshold=0.5)

    @property
    def alpha(self) -> torch.Tensor:
        return F.softplus(self.beta)

    def _flags(self) -> int:
        f = 0
        if self.use_sigmoid_mask:
            f |= _lib.SIGMOID_MASK
        if self.sam_cam_fusion in ("add", "concat"):  # the spatial gate is computed from x itself (parallel CBAM)
            f |= _lib.SAMCAM_ADD
        if self.mga_pyramid_fusion == "multiply":
            f |= _lib.PYRAMID_MULTIPLY
        if os.getenv("MGA_FORCE_SPLIT", ""):  # one kernel per phase instead of the cluster-per-sample kernels
            f |= _lib.FORCE_SPLIT
        if os.getenv("MGA_NO_PERSIST", ""):  # cluster-per-sample kernels instead of the persistent shared-memory-resident ones
            f |= _lib.NO_PERSIST
        return f

    def forward(self, x: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
        if isinstance(x, (list, tuple)):
            assert len(x) == 2, "MaskGuidedCBAM expects [feature, mask]"
            feat, mask = x
        else:
            feat, mask = x, None
        assert isinstance(feat, torch.Tensor) and feat.dim() == 4
        if feat.device.type == "cpu":
            if in_shape_probe():
                return torch.zeros_like(feat)  # shape-only answer (see module docstring); no math runs on the CPU
            raise RuntimeError(
                "mga_yolo_b200: the mask-guided CBAM path runs on CUDA tensors only (no CPU fallback); move the model and its "
                "inputs to a GPU.  Shape probes of a graph builder go through mga_yolo_b200.shape_probe() (install() does that "
                "for the reference's DetectionModel.__init__) or through meta tensors.")
        flags = self._flags()
        if mask is not None:
            if os.getenv("MGA_PROB_MODE", False) and hasattr(self, "gater"):
                if self.gater.is_deterministic():
                    flags |= _lib.GATE_CLAMP  # clamp fused into the kernels
                    if self.gater.p_min > 0:
                        mask = mask.float().clamp_min(self.gater.p_min)
                else:
                    mask = self.gater.sample(mask)
            if not mask.is_floating_point():
                mask = mask.to(torch.float32)  # binary uint8/bool maps (use_sigmoid_mask=False)
            elif mask.dtype == torch.float64:
                mask = mask.float()
        lin1, lin2 = self.cam_mlp[0], self.cam_mlp[2]
        if "concat" in (self.sam_cam_fusion, self.mga_pyramid_fusion):
            return self._forward_concat(feat, mask, flags)
        return ops.mask_guided_cbam(
            feat, mask, lin1.weight, lin1.bias, lin2.weight, lin2.bias, self.sam_conv.weight, self.beta,
            flags=flags, tiny_mask_thr=self.tiny_thr, eps=self.eps,
        )

    def _forward_concat(self, feat: torch.Tensor, mask: Optional[torch.Tensor], flags: int) -> torch.Tensor:
        """concat fusion modes: gates from the CUDA op, the 2C->C 1x1 convolutions from the library (cuDNN / cuBLAS)."""
        lin1, lin2 = self.cam_mlp[0], self.cam_mlp[2]
        dt = feat.dtype
        if self.sam_cam_fusion == "multiply":  # R = x*s*a with a from x*s: the fused op with alpha == 1 and no identity skip
            one = torch.full((), _BETA_ALPHA_ONE, dtype=torch.float32, device=feat.device)
            R = ops.mask_guided_cbam(feat, mask, lin1.weight, lin1.bias, lin2.weight, lin2.bias, self.sam_conv.weight, one,
                                     flags=(flags | _lib.PYRAMID_MULTIPLY), tiny_mask_thr=self.tiny_thr, eps=self.eps)
        else:
            if (self.sam_cam_fusion == "concat" and self.mga_pyramid_fusion in ("add", "multiply") and next_ops.concat_fused_supported(feat)
                    and not os.getenv("MGA_CONCAT_LIBRARY", "")):
                # 16-bit features with C % 128 == 0: gates op + ONE tcgen05 kernel that does the (virtual) concat, both 1x1 convolutions,
                # the spatial gate and the pyramid fusion (csrc/cbam_concat.cuh), as one autograd node (closed-form backward)
                return next_ops.concat_block(feat, mask, lin1.weight, lin1.bias, lin2.weight, lin2.bias, self.sam_conv.weight,
                                             self.fuse_sam_cam.weight, self.fuse_sam_cam.bias, self.beta,
                                             flags=flags & ~_lib.PYRAMID_MULTIPLY, tiny_mask_thr=self.tiny_thr, eps=self.eps,
                                             pyramid_multiply=self.mga_pyramid_fusion == "multiply")
            s, a = ops.cbam_gates(feat, mask, lin1.weight, lin1.bias, lin2.weight, lin2.bias, self.sam_conv.weight,
                                  flags=flags & ~_lib.PYRAMID_MULTIPLY, tiny_mask_thr=self.tiny_thr, eps=self.eps)
            xs = feat * s.to(dt)[:, :, None, None]
            xa = feat * a.to(dt)
            if self.sam_cam_fusion == "concat":
                R = F.conv2d(torch.cat([xs, xa], dim=1), self.fuse_sam_cam.weight.to(dt), self.fuse_sam_cam.bias.to(dt))
            else:
                R = xs + xa
        if self.mga_pyramid_fusion == "concat":
            return F.conv2d(torch.cat([feat, R], dim=1), self.fuse_pyramid.weight.to(dt), self.fuse_pyramid.bias.to(dt))
        alpha = self.alpha.to(dt)
        if self.mga_pyramid_fusion == "multiply":
            return alpha * R
        return feat + alpha * (R - feat)

    def extra_repr(self) -> str:  # pragma: no cover
        return (f"channels={self.C}, r={self.r}, spatial_k={self.k}, use_sigmoid_mask={self.use_sigmoid_mask}, "
                f"tiny_mask_thr={self.tiny_thr}, eps={self.eps}, sam_cam_fusion={self.sam_cam_fusion}, "
                f"mga_pyramid_fusion={self.mga_pyramid_fusion}, alpha={self.alpha.item():.4f}")


class MaskCBAM(MaskGuidedCBAM):
    """Name used by the reference's model YAML / parse_model (configs/models/yolov8_cbam.yaml:67-72,
    ultralytics/nn/tasks.py:1733-1739)."""
