"""ctypes binding of libmga_cbam.so (C ABI declared in include/mga_cbam.h).

There is deliberately no CPU implementation behind these symbols: if the shared
library is missing or does not export what the header declares, every op raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os as _os

LIB_PATH = Path(__file__).resolve().parent / _os.environ.get("MGA_LIBNAME", "libmga_cbam.so")
ABI_VERSION = 3

# enums of include/mga_cbam.h
F32, BF16, F16, U8 = 0, 1, 2, 3
HAS_MASK, SIGMOID_MASK, GATE_CLAMP = 1 << 0, 1 << 1, 1 << 2
SAMCAM_ADD, PYRAMID_MULTIPLY, FORCE_SPLIT, GATES_ONLY, NO_SAVE, NO_PERSIST = 1 << 4, 1 << 6, 1 << 8, 1 << 10, 1 << 12, 1 << 14
DS_NEAREST, DS_AREA, DS_MAXPOOL, DS_AVGPOOL, DS_AREA_RAW = 0, 1, 2, 3, 4

EXPORTS = (
    "mga_abi_version", "mga_last_error", "mga_cbam_workspace", "mga_cbam_forward", "mga_cbam_backward",
    "mga_cbam_ctx_view", "mga_mask_downsample", "mga_masks_multi", "mga_masks_multi_ws", "mga_cbam_plan", "mga_cbam_gates_forward", "mga_cbam_gates_backward", "mga_cbam_gates_backward_acc", "mga_launch_count", "mga_profile_enable", "mga_profile_count",
    "mga_profile_read", "mga_eca_workspace", "mga_eca_forward", "mga_eca_backward", "mga_head_tail_forward", "mga_head_tail_backward",
    "mga_gate_sample_forward", "mga_gate_sample_backward", "mga_collate_masks", "mga_cbam_concat_forward", "mga_cbam_concat_backward_elem",
    "mga_cbam_concat_wgrad_reduce", "mga_cbam_concat_backward_dx", "mga_spade_forward", "mga_spade_backward",
)


class Desc(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("hidden", C.c_int32), ("ksize", C.c_int32), ("dtype", C.c_int32), ("mask_dtype", C.c_int32),
        ("flags", C.c_int32), ("tiny_mask_thr", C.c_float), ("eps", C.c_float),
    ]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1", "b1", "w2", "b2", "wsam", "beta")]


Grads = Params  # same layout: six float pointers


class PlanInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("path", "cluster_size", "rows_per_cta", "threads", "smem_bytes", "launches")]

_lib = None


class MgaLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library once; fail loudly if it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise MgaLibraryError(
            f"{LIB_PATH} is missing. Build it with `python -m mga_yolo_b200.build` (needs nvcc, sm_100a). "
            "mga_yolo_b200 has no CPU or PyTorch fallback for the mask-guided CBAM path."
        )
    lib = C.CDLL(str(LIB_PATH))
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise MgaLibraryError(f"{LIB_PATH} does not export {missing}; rebuild it")
    lib.mga_abi_version.restype = C.c_int
    lib.mga_last_error.restype = C.c_char_p
    lib.mga_cbam_workspace.argtypes = [C.POINTER(Desc), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.mga_cbam_forward.argtypes = [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    lib.mga_cbam_backward.argtypes = [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(Grads), C.c_void_p, C.c_void_p]
    lib.mga_cbam_gates_forward.argtypes = [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mga_cbam_gates_backward.argtypes = [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params), C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.POINTER(Grads), C.c_void_p, C.c_void_p]
    lib.mga_cbam_gates_backward_acc.argtypes = [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Params),
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Grads), C.c_void_p, C.c_void_p]
    lib.mga_cbam_gates_forward.restype = C.c_int
    lib.mga_cbam_gates_backward.restype = C.c_int
    lib.mga_cbam_gates_backward_acc.restype = C.c_int
    lib.mga_cbam_ctx_view.argtypes = [C.POINTER(Desc), C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.mga_mask_downsample.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_void_p]
    lib.mga_masks_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                    C.c_int32, C.c_int32, C.c_void_p]
    lib.mga_masks_multi.restype = C.c_int
    lib.mga_masks_multi_ws.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_float, C.c_int32, C.c_int32, C.c_void_p]
    lib.mga_masks_multi_ws.restype = C.c_int
    lib.mga_cbam_plan.argtypes = [C.POINTER(Desc), C.c_int, C.POINTER(PlanInfo)]
    lib.mga_cbam_plan.restype = C.c_int
    lib.mga_launch_count.restype = C.c_ulonglong
    lib.mga_profile_enable.argtypes = [C.c_int]
    lib.mga_profile_read.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    lib.mga_eca_workspace.argtypes = [C.POINTER(Desc), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.mga_eca_forward.argtypes = [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mga_eca_backward.argtypes = [C.POINTER(Desc)] + [C.c_void_p] * 11
    lib.mga_head_tail_forward.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 5 + [C.c_void_p]
    lib.mga_head_tail_backward.argtypes = [C.c_void_p] * 6 + [C.c_int32] * 5 + [C.c_void_p]
    lib.mga_gate_sample_forward.argtypes = [C.c_void_p] * 5 + [C.c_size_t, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_uint64, C.c_void_p]
    lib.mga_gate_sample_backward.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_float, C.c_float, C.c_void_p]
    lib.mga_collate_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.mga_cbam_concat_forward.argtypes = [C.POINTER(Desc)] + [C.c_void_p] * 9
    lib.mga_cbam_concat_forward.restype = C.c_int
    lib.mga_cbam_concat_backward_elem.argtypes = [C.POINTER(Desc)] + [C.c_void_p] * 14
    lib.mga_cbam_concat_backward_elem.restype = C.c_int
    lib.mga_cbam_concat_wgrad_reduce.argtypes = [C.POINTER(Desc), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mga_cbam_concat_wgrad_reduce.restype = C.c_int
    lib.mga_cbam_concat_backward_dx.argtypes = [C.POINTER(Desc)] + [C.c_void_p] * 15
    lib.mga_cbam_concat_backward_dx.restype = C.c_int
    lib.mga_spade_forward.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_float, C.c_int32, C.c_int32, C.c_void_p]
    lib.mga_spade_forward.restype = C.c_int
    lib.mga_spade_backward.argtypes = [C.c_void_p] * 6 + [C.c_int32] * 6 + [C.c_void_p]
    lib.mga_spade_backward.restype = C.c_int
    for fn in (lib.mga_eca_workspace, lib.mga_eca_forward, lib.mga_eca_backward, lib.mga_head_tail_forward, lib.mga_head_tail_backward,
               lib.mga_gate_sample_forward, lib.mga_gate_sample_backward, lib.mga_collate_masks):
        fn.restype = C.c_int
    for fn in (lib.mga_cbam_workspace, lib.mga_cbam_forward, lib.mga_cbam_backward, lib.mga_cbam_ctx_view, lib.mga_mask_downsample):
        fn.restype = C.c_int
    if lib.mga_abi_version() != ABI_VERSION:
        raise MgaLibraryError(f"ABI mismatch: library {lib.mga_abi_version()} vs binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mga_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
