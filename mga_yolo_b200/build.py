"""Build libmga_cbam.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

    python -m mga_yolo_b200.build          # or: __graft_entry__.build()

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libmga_cbam.so"
LIB_TUNING = PKG / "libmga_cbam_tuning.so"  # same sources + -DMGA_TUNING: MGA_CL_* environment knobs and the phase timeline (tools/, geometry tests)
SOURCES = ("mga_cbam.cu", "mask_ops.cu", "spade_ops.cu")
ARCH = ("-gencode", "arch=compute_100a,code=sm_100a")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _stale(lib: Path = LIB) -> bool:
    if not lib.exists():
        return True
    t = lib.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, tuning: bool = False) -> Path:
    lib = LIB_TUNING if tuning else LIB
    if not force and not _stale(lib):
        return lib
    nvcc = _nvcc()
    objdir = PKG / "build" / ("tuning" if tuning else "product")
    objdir.mkdir(parents=True, exist_ok=True)
    common = [nvcc, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", str(INCLUDE), "-I", str(CSRC)]
    if verbose:
        common += ["-Xptxas", "-v"]
    if tuning:
        common += ["-DMGA_TUNING"]
        common += [f"-D{d}" for d in os.environ.get("MGA_DEFINES", "").split() if d]

    def compile_one(src: str) -> Path:
        obj = objdir / (src + ".o")
        cmd = common + ["-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = lib.with_suffix(".so.tmp")
    r = subprocess.run([nvcc, *ARCH, "-shared", "-o", str(tmp), *map(str, objs), "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, lib)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tuning="--tuning" in sys.argv))
