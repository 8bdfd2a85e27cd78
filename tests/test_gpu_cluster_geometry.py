"""The cluster-per-sample kernels with FORCED cluster sizes on shapes whose rows do not divide evenly: ragged last rank,
halo rows that come from more than one neighbour, 16-bit units that span image rows (W = 20), cluster size 16.
Every case runs in a child process because the tuning overrides are read once per process.  Oracle: oracle/cbam_oracle.py
(fp64); tolerances as in test_gpu_cbam.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

CASES = [
    # B, C,  H,  W, dtype,      scf,        pyr,        CS_F, CS_B
    (3, 48, 22, 24, "float32", "multiply", "add", 4, 4),     # rows 6,6,6,4
    (2, 40, 20, 20, "float32", "multiply", "add", 8, 8),     # rows 3 x6 + 2: halo from two ranks away never needed, last rank short
    (2, 32, 16, 16, "float32", "add", "multiply", 16, 16),   # one row per CTA: the 3 halo rows come from three different ranks
    (2, 24, 80, 80, "float32", "multiply", "add", 16, 16),   # cluster of 16, 5 rows each
    (2, 64, 20, 20, "bfloat16", "multiply", "add", 4, 2),    # W = 20 with 8-element units: rows per CTA must be even
    (2, 64, 40, 40, "float16", "add", "add", 8, 4),
    (5, 128, 40, 40, "float32", "multiply", "multiply", 2, 8),
    # default geometry (0 = not forced) on wide levels: YOLOv8x P5 at 1280 (C = 768, hidden = 48: 177 / 205 KB of shared memory,
    # one CTA per SM) and YOLOv8s/m P5 (C = 512)
    (2, 768, 40, 40, "bfloat16", "multiply", "add", 0, 0),
    (2, 768, 40, 40, "float32", "multiply", "add", 0, 0),   # 4.9 MB per sample: too large for the cluster path -> per-phase kernels
    (3, 512, 20, 20, "float32", "add", "add", 0, 0),
    (2, 512, 20, 20, "float16", "multiply", "multiply", 0, 0),
]


# the other input modes of the block on multi-CTA clusters: no mask (vanilla CBAM, masked_cbam.py:90-91,107-108,137-138), a raw
# {0,1} mask with use_sigmoid_mask=False, a mask that needs no gradient (ground-truth masks)
CASES += [
    (2, 48, 24, 24, "float32", "multiply", "add", 4, 4, "nomask"),
    (2, 48, 24, 24, "float32", "multiply", "add", 4, 4, "rawmask"),
    (2, 48, 24, 24, "float32", "add", "add", 4, 2, "nograd_mask"),
    (2, 64, 40, 40, "bfloat16", "multiply", "add", 2, 2, "rawmask"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[1]}x{c[2]}x{c[3]}-{c[4]}-cs{c[7]}/{c[8]}" + (f"-{c[9]}" if len(c) > 9 else ""))
def test_forced_cluster_sizes_match_oracle(case):
    B, C, H, W, dt, scf, pyr, csf, csb = case[:9]
    variant = case[9] if len(case) > 9 else "default"
    # forcing a cluster size is a tuning knob: only the -DMGA_TUNING build of the library (same kernels) reads MGA_CL_*
    env = dict(os.environ, MGA_CL_CS_F=str(csf), MGA_CL_CS_B=str(csb), MGA_CL_DEBUG="1", PYTHONPATH=str(ROOT), MGA_LIBNAME="libmga_cbam_tuning.so")
    for k in ("MGA_FORCE_SPLIT", "MGA_CL"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "tests._cluster_case", str(B), str(C), str(H), str(W), dt, scf, pyr, variant], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    if csf and csb:  # forced sizes must really have gone through the cluster kernels (default geometry may pick the per-phase path)
        assert "[mga] cluster fwd" in r.stderr and "[mga] cluster bwd" in r.stderr, "the cluster path did not run:\n" + r.stderr
