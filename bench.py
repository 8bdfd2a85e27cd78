#!/usr/bin/env python
"""bench.py -- MGA-CBAM fwd+bwd throughput on synthetic YOLOv8 neck-shaped tensors.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg3|cfg5]

One "step" = forward + backward of the mask-guided CBAM block on all three pyramid levels
(P3/P4/P5) for one batch.  Default workload = BASELINE.json configs[1]: YOLOv8n shapes
(64,80,80)/(128,40,40)/(256,20,20), batch 64, fp32.

Numbers on the JSON line
  value          algorithmic GB/s of the whole step, inputs resident in HBM, CUDA events on the
                 launching stream, max over ranks.  algorithmic bytes = (5N + 3BS)*e per level
                 (SURVEY.md section 8d): read x, read mask, write out; read x, read g, write dx,
                 read mask, write dmask.
  images_per_sec batch / step time (whole job).
  e2e            the same metric through the public nn.Module API (MaskGuidedCBAM + autograd) with
                 HOST buffers: every step copies x, mask and grad_out from pinned host memory and
                 reads the parameter gradients back.
  roofline       dominant kernel: algorithmic bytes of that kernel / its CUDA-event duration,
                 measured in an instrumented pass right after the timed region (per-kernel events
                 cannot be recorded inside a CUDA graph).
  cpu_baseline   the reference's own MaskCBAM class (oracle/_ref, laid out by oracle/build_ref.py;
                 kind "reference") -- or, when oracle/_ref is absent, the oracle port (kind "port") --
                 on the host cores, bounded sample of the same workload.
  gpu_eager_baseline  the reference class .cuda() (eager PyTorch kernels) on the same B200, same tensors:
                 the like-for-like GPU comparator of SURVEY.md section 8d.
  workloads      the other single-GPU BASELINE configs (cfg3 = configs[2] shapes, cfg5 = configs[4]
                 shapes), same step, same timing rules, fewer timed steps.
Multi-GPU: one process per GPU (torchrun).  --scaling weak (default): every rank runs the full per-GPU
batch; --scaling strong: the workload's GLOBAL batch is sharded (BASELINE configs[2]: 256 -> 128/64/32).
The only collective is one NCCL all-reduce of the flat weight-gradient buffer.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch

WORKLOADS = {
    # name: (levels (C,H,W), per-GPU batch, dtype, description)
    "cfg2": ([(64, 80, 80), (128, 40, 40), (256, 20, 20)], 64, "float32",
             "BASELINE configs[1]: MGA-CBAM module-only fwd+bwd, YOLOv8n P3/P4/P5, batch 64 fp32"),
    "cfg3": ([(128, 80, 80), (256, 40, 40), (512, 20, 20)], 256, "bfloat16",
             "BASELINE configs[2] shapes: YOLOv8s P3/P4/P5, batch 256 bf16 (module-only)"),
    "cfg5": ([(384, 160, 160), (768, 80, 80), (768, 40, 40)], 8, "bfloat16",
             "BASELINE configs[4] shapes: YOLOv8x imgsz 1280, batch 8 per GPU bf16 (module-only)"),
}
DT = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


def algorithmic_bytes(levels, B, esize, mask_esize=4):
    tot = 0
    for (Cc, H, W) in levels:
        N, S = B * Cc * H * W, H * W
        tot += 5 * N * esize + 3 * B * S * mask_esize
    return tot


# per-kernel algorithmic bytes (what one launch must move at minimum), in units of N*e / B*S*4
KERNEL_BYTES = {
    "cam_pool": lambda N, BS, e: N * e + BS * 4,          # read x, read m
    "sam_reduce": lambda N, BS, e: N * e,                 # read x
    "rescale": lambda N, BS, e: 2 * N * e + BS * 4,       # read x, write out, read a
    "bwd_reduce1": lambda N, BS, e: 2 * N * e + BS * 4,   # read x, g, a
    "bwd_reduce2": lambda N, BS, e: N * e + 3 * BS * 4,   # read x, dcat0/1, idx
    "bwd_dx": lambda N, BS, e: 3 * N * e + 6 * BS * 4,    # read x, g, write dx (+ planes, dmask)
    "fused_fwd": lambda N, BS, e: 2 * N * e + 2 * BS * 4,
    "flow_fwd": lambda N, BS, e: 2 * N * e + BS * 4,      # the whole forward: read x, mask; write out
    "flow_bwd": lambda N, BS, e: 3 * N * e + 2 * BS * 4,  # the whole backward: read x, g, mask; write dx, dmask
    "fused_bwd": lambda N, BS, e: 3 * N * e + 3 * BS * 4,
    "cl_fwd": lambda N, BS, e: 2 * N * e + BS * 4,        # cluster-per-sample forward: read x, mask; write out
    "cl_bwd": lambda N, BS, e: 3 * N * e + 2 * BS * 4,    # cluster-per-sample backward: read x, g, mask; write dx, dmask
}


# --------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the GPU phase runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self._stop = [], threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.time(), sm, util, rs, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop.set()

    def summary(self, t0: float, t1: float):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + getattr(self, "err", "?")}
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        win = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        if not win:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": 0}
        reasons = set()
        for s in win:
            for bit, nm in names.items():
                if s[3] & bit:
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(s[1] for s in win), "sm_max_mhz": self.max_sm, "reasons": sorted(reasons),
                "samples": len(win), "power_w_max": round(max(s[4] for s in win), 1)}


# --------------------------------------------------------------------------- parameters / CPU arm
def make_params(Cc, r=16, k=7, beta=0.0):
    """Deterministic parameter values: the block's own default init (== the reference's init order, masked_cbam.py:53-64)
    under torch.manual_seed(C).  Returned in state_dict order of the flat gradient buffer: w1, b1, w2, b2, wsam, beta."""
    from mga_yolo_b200 import MaskGuidedCBAM

    torch.manual_seed(Cc)
    m = MaskGuidedCBAM(Cc, r=r, spatial_k=k)
    with torch.no_grad():
        m.beta.fill_(beta)
    return [t.detach().clone() for t in (m.cam_mlp[0].weight, m.cam_mlp[0].bias, m.cam_mlp[2].weight, m.cam_mlp[2].bias, m.sam_conv.weight, m.beta)]


def reference_available():
    try:
        from oracle import build_ref
        return build_ref.available()
    except Exception:
        return False


def cpu_step_fn(levels, B, dtype, sam_cam, threads):
    """fwd+bwd of the three levels on CPU tensors.  kind "reference": the reference's MaskCBAM class itself (oracle/_ref) with
    torch autograd -- only for the reference-equivalent fusion mode; kind "port": oracle/cbam_oracle.py."""
    torch.set_num_threads(threads)
    gen = torch.Generator().manual_seed(0)
    use_ref = reference_available() and sam_cam == "multiply"
    work = []
    if use_ref:
        from oracle import build_ref

        ref = build_ref.load()
    else:
        from oracle import cbam_oracle as co
    for (Cc, H, W) in levels:
        x = torch.randn(B, Cc, H, W, generator=gen).to(dtype).float()
        mk = torch.randn(B, 1, H, W, generator=gen)
        g = torch.randn(B, Cc, H, W, generator=gen)
        w1, b1, w2, b2, wsam, beta = make_params(Cc)
        if use_ref:
            mod = ref.MaskCBAM(Cc)
            mod.load_state_dict({"cam_mlp.0.weight": w1, "cam_mlp.0.bias": b1, "cam_mlp.2.weight": w2, "cam_mlp.2.bias": b2,
                                 "sam_conv.weight": wsam, "beta": beta})
            work.append((x, mk, g, mod))
        else:
            leaves = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2, wsam, beta)]
            work.append((x, mk, g, leaves))

    def step():
        for x, mk, g, obj in work:
            xi = x.clone().requires_grad_(True)
            mi = mk.clone().requires_grad_(True)
            if use_ref:
                obj([xi, mi]).backward(g)
                obj.zero_grad(set_to_none=True)
            else:
                out = co.cbam_forward_autograd(xi, mi, co.CbamParams(*obj), sam_cam_fusion=sam_cam)
                out.backward(g)
                for leaf in obj:
                    leaf.grad = None

    return step, ("reference" if use_ref else "port")


def time_cpu(levels, B, dtype, sam_cam, steps, warmup):
    threads = os.cpu_count() or 1
    step, kind = cpu_step_fn(levels, B, dtype, sam_cam, threads)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), threads, kind


# --------------------------------------------------------------------------- GPU arm (C ABI, resident inputs)
class LevelPlan:
    """Pre-allocated buffers + descriptor for one pyramid level, driven straight through the C ABI."""

    def __init__(self, lib_mod, Cc, H, W, B, dtype, flags, dev, seed, grads_flat, goff):
        from mga_yolo_b200 import _lib

        self.lib = _lib.load()
        self._lib = _lib
        self.shape = (B, Cc, H, W)
        gen = torch.Generator(device=dev).manual_seed(seed)
        self.x = torch.randn(B, Cc, H, W, generator=gen, device=dev).to(dtype)
        self.mask = torch.randn(B, 1, H, W, generator=gen, device=dev)
        self.g = torch.randn(B, Cc, H, W, generator=gen, device=dev).to(dtype)
        self.out = torch.empty_like(self.x)
        self.dx = torch.empty_like(self.x)
        self.dmask = torch.empty_like(self.mask)
        self.params = [t.to(dev).contiguous() for t in make_params(Cc)]
        hidden = self.params[0].shape[0]
        dt_code = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}[dtype]
        self.desc = _lib.Desc(B, Cc, H, W, hidden, 7, dt_code, _lib.F32, flags | _lib.HAS_MASK | _lib.SIGMOID_MASK, 1e-4, 1e-6)
        cb, sb = C.c_size_t(0), C.c_size_t(0)
        _lib.check(self.lib.mga_cbam_workspace(C.byref(self.desc), C.byref(cb), C.byref(sb)), "workspace")
        self.ctx = torch.empty(cb.value, dtype=torch.uint8, device=dev)
        self.scratch = torch.empty(sb.value, dtype=torch.uint8, device=dev)
        self.prm = _lib.Params(*(t.data_ptr() for t in self.params))
        sizes = [t.numel() for t in self.params]
        offs, o = [], goff
        for n in sizes:
            offs.append(o)
            o += n
        self.grad_end = o
        self.gp = _lib.Grads(*(grads_flat.data_ptr() + 4 * v for v in offs))

    @staticmethod
    def n_params(Cc, r=16, k=7):
        h = max(1, Cc // r)
        return h * Cc + h + Cc * h + Cc + 3 * k * k + 1

    def fwd(self, stream):
        rc = self.lib.mga_cbam_forward(C.byref(self.desc), self.x.data_ptr(), self.mask.data_ptr(), C.byref(self.prm),
                                       self.out.data_ptr(), self.ctx.data_ptr(), self.scratch.data_ptr(), stream)
        self._lib.check(rc, "mga_cbam_forward")

    def bwd(self, stream):
        rc = self.lib.mga_cbam_backward(C.byref(self.desc), self.x.data_ptr(), self.mask.data_ptr(), self.g.data_ptr(),
                                        C.byref(self.prm), self.ctx.data_ptr(), self.dx.data_ptr(), self.dmask.data_ptr(),
                                        C.byref(self.gp), self.scratch.data_ptr(), stream)
        self._lib.check(rc, "mga_cbam_backward")


def build_plans(levels, B, dtype, flags, dev, nsets):
    total = sum(LevelPlan.n_params(c) for c, _, _ in levels)
    sets = []
    for si in range(nsets):
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        plans, off = [], 0
        for li, (Cc, H, W) in enumerate(levels):
            pl = LevelPlan(None, Cc, H, W, B, dtype, flags, dev, 1000 * si + li, flat, off)
            off = pl.grad_end
            plans.append(pl)
        sets.append((plans, flat))
    return sets


def run_step(plans, stream):
    """Serial step on one stream: forward of P3,P4,P5 then backward of P5,P4,P3 (training order)."""
    for pl in plans:
        pl.fwd(stream)
    for pl in reversed(plans):
        pl.bwd(stream)


def run_step_streams(plans, main, sides):
    """Same step with one stream per pyramid level (the three levels are independent): all forwards
    run concurrently, join, then all backwards run concurrently, join."""
    streams = [main] + list(sides)
    for phase in ("fwd", "bwd"):
        for s in sides:
            s.wait_stream(main)
        for pl, s in zip(plans, streams):
            getattr(pl, phase)(s.cuda_stream)
        for s in sides:
            main.wait_stream(s)


def measure(args, lib, dev, world, levels, B, dtype, flags, steps, warmup, one_stream=False, instrument=True):
    """Time `steps` fwd+bwd steps of one workload (CUDA events on the launching stream, max over ranks) and, in an
    instrumented pass right after, every kernel of the step (events around each launch, one stream)."""
    import torch.distributed as dist

    from mga_yolo_b200 import _lib

    sets = build_plans(levels, B, dtype, flags, dev, nsets=2)
    stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream
    # the side streams (one per extra level) run at HIGH priority: under torchrun the asynchronous NCCL all-reduce of the previous step
    # then never takes CTA slots from the cluster kernels of the current one (+11 us per step at N >= 2)
    sides = [torch.cuda.Stream(dev, priority=-1) for _ in levels[1:]] if not (args.one_stream or one_stream) else []

    def step_on(plans, main):
        if sides:
            run_step_streams(plans, main, sides)
        else:
            run_step(plans, main.cuda_stream)

    # warm-up (un-graphed) also gives launches per step
    torch.cuda.synchronize(dev)
    n0 = lib.mga_launch_count()
    run_step(sets[0][0], sptr)
    launches_per_step = lib.mga_launch_count() - n0
    run_step(sets[1][0], sptr)
    torch.cuda.synchronize(dev)
    graphs = None
    if not args.no_graph:
        try:
            graphs = []
            for plans, _flat in sets:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step_on(plans, torch.cuda.current_stream(dev))
                graphs.append(g)
        except Exception as e:  # pragma: no cover
            print(f"[bench] CUDA graph capture failed ({e}); timing direct launches", file=sys.stderr)
            graphs = None

    pending = [None, None]  # all-reduce in flight on each buffer set

    def one(i):
        if pending[i & 1] is not None:  # this set's previous gradients must have been reduced before they are overwritten
            pending[i & 1].wait()
            pending[i & 1] = None
        if graphs is not None:
            graphs[i & 1].replay()
        else:
            step_on(sets[i & 1][0], stream)
        if world > 1:
            # the only collective: the flat weight-gradient buffer; asynchronous (NCCL's stream waits for this step's kernels,
            # the next step's kernels -- on the other buffer set -- do not wait for NCCL), like DDP's bucket overlap
            pending[i & 1] = dist.all_reduce(sets[i & 1][1], async_op=True)

    def drain():
        for k in (0, 1):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    for i in range(warmup):
        one(i)
    drain()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(steps):
        one(i)
    drain()  # every all-reduce of the timed steps is inside the timed region
    e1.record()
    torch.cuda.synchronize(dev)
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())

    klist = []
    if instrument:
        # per-kernel CUDA events (same buffers, direct launches, one stream); every record is tagged with the level whose call produced it
        lib.mga_profile_enable(1)
        reps = 5
        tags = []
        for i in range(reps):
            plans = sets[i & 1][0]
            for phase, order in (("fwd", list(enumerate(plans))), ("bwd", list(enumerate(plans))[::-1])):
                for li, pl in order:
                    n0 = lib.mga_profile_count()
                    getattr(pl, phase)(sptr)
                    tags += [(i, li, phase)] * (lib.mga_profile_count() - n0)
        torch.cuda.synchronize(dev)
        name, val = C.c_char_p(), C.c_float()
        per_kernel = {}
        order_keys = []
        for idx in range(lib.mga_profile_count()):
            _lib.check(lib.mga_profile_read(idx, C.byref(name), C.byref(val)), "profile_read")
            rep, li, phase = tags[idx]
            if rep == 0:
                continue  # skip the first repetition
            key = (phase, li, name.value.decode())
            if key not in per_kernel:
                order_keys.append(key)
            per_kernel.setdefault(key, []).append(val.value)
        lib.mga_profile_enable(0)
        klist = [{"kernel": nm, "li": li, "phase": phase, "ms": sum(per_kernel[(phase, li, nm)]) / len(per_kernel[(phase, li, nm)])}
                 for (phase, li, nm) in order_keys]
    return {"ms": ms, "launches_per_step": int(launches_per_step), "graph": graphs is not None, "kernels": klist,
            "streams": 1 + len(sides), "wall": (t_wall0, t_wall1), "sets": sets}


def annotate_kernels(klist, levels, B, esize, peak):
    """Algorithmic bytes / GB/s per kernel record; returns the dominant (longest) one."""
    best = None
    for k in klist:
        Cc, H, W = levels[k["li"]]
        N, BS = B * Cc * H * W, B * H * W
        fn = KERNEL_BYTES.get(k["kernel"])
        k["level"] = f"P{3 + k['li']}"
        if fn is None:
            continue
        k["alg_bytes"] = fn(N, BS, esize)
        k["gbps"] = k["alg_bytes"] / (k["ms"] * 1e-3) / 1e9
        k["frac"] = k["gbps"] / peak
        if best is None or k["ms"] > best["ms"]:
            best = k
    return best


def gpu_arm(args, rank, world, local_rank):
    from mga_yolo_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    levels, Bglobal, dtname, desc_txt = WORKLOADS[args.workload]
    if args.batch:
        Bglobal = args.batch
    strong = args.scaling == "strong"
    if strong and Bglobal % world:
        raise SystemExit(f"--scaling strong: the global batch {Bglobal} does not divide over {world} ranks")
    B = Bglobal // world if strong else Bglobal
    dtype = DT[dtname]
    esize = torch.empty((), dtype=dtype).element_size()

    def flags_of(scf, split=False):
        f = _lib.SAMCAM_ADD if scf == "add" else 0
        return f | (_lib.FORCE_SPLIT if (args.force_split or split) else 0)

    sampler = ClockSampler(local_rank)
    sampler.start()
    alg_bytes = algorithmic_bytes(levels, B, esize)
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (of fallback)"

    main = measure(args, lib, dev, world, levels, B, dtype, flags_of(args.sam_cam_fusion), args.steps, args.warmup)
    t0w, t1w = main["wall"]
    # keep the GPU busy long enough for the clock sampler when the timed region is short
    if t1w - t0w < 1.0:
        stream = torch.cuda.current_stream(dev).cuda_stream
        tb = time.time()
        while time.time() - tb < 1.0:
            for _ in range(20):
                run_step(main["sets"][0][0], stream)
            torch.cuda.synchronize(dev)
        t1w = time.time()
    main["sets"] = None
    short = (max(10, args.steps // 2), max(3, args.warmup))
    other = "add" if args.sam_cam_fusion == "multiply" else "multiply"
    variant = None if args.no_variant else measure(args, lib, dev, world, levels, B, dtype, flags_of(other), *short, instrument=False)
    # the other launch path of the library (one kernel per phase), same step, same timing rules
    variant_split = None if (args.no_variant or args.force_split) else \
        measure(args, lib, dev, world, levels, B, dtype, flags_of(args.sam_cam_fusion, split=True), *short, instrument=False)
    # the same step issued on ONE stream (the order an nn.Module caller gets), graph replay
    one_stream = None if (args.no_variant or args.one_stream) else \
        measure(args, lib, dev, world, levels, B, dtype, flags_of(args.sam_cam_fusion), *short, one_stream=True, instrument=False)
    for v in (variant, variant_split, one_stream):
        if v is not None:
            v["sets"] = None
    torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel
    klist = main["kernels"]
    best = annotate_kernels(klist, levels, B, esize, peak)
    roofline = None
    traffic = None
    tpath = ROOT / "profiles" / "traffic.json"
    if best and tpath.exists():  # dram bytes per launch from the committed ncu --set full capture of the same kernel/shape
        traffic = json.loads(tpath.read_text()).get(f"{best['kernel']}[{best['level']}]") if (args.workload == "cfg2" and not strong and not args.batch) else None
    if best:
        roofline = {"bound": "hbm", "kernel": f"{best['kernel']}[{best['level']}]", "achieved": round(best["gbps"], 1), "peak": peak,
                    "unit": "GB/s", "frac": round(best["gbps"] / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "launch_ms": round(best["ms"], 5), "alg_bytes_per_launch": best["alg_bytes"],
                    "step_frac": round(alg_bytes / (main["ms"] * 1e-3) / 1e9 / peak, 4),
                    "step_frac_of_8TBps_nominal": round(alg_bytes / (main["ms"] * 1e-3) / 1e9 / 8000.0, 4)}

    # ---- e2e: public module API, host buffers
    e2e = None if args.no_e2e else e2e_module(args, dev, local_rank, levels, B, dtype, world, alg_bytes)

    sampler.stop()
    clocks = sampler.summary(t0w, t1w)
    value = world * alg_bytes / (main["ms"] * 1e-3) / 1e9
    nparams = sum(LevelPlan.n_params(c) for c, _, _ in levels)
    line = {
        "metric": "mga_cbam_fwd_bwd_algorithmic_GBps", "value": round(value, 1), "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(main["ms"], 5), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": {"float32": "f32", "bfloat16": "bf16", "float16": "f16"}[dtname],
        "data": "synthetic (x, grad_out ~ N(0,1); mask logits ~ N(0,1); the block's reference-order default parameter init under torch.manual_seed(C))",
        "images_per_sec": round(world * B / (main["ms"] * 1e-3), 1),
        "config": {"workload": desc_txt, "levels_CHW": levels, "batch_per_gpu": B, "global_batch": B * world,
                   "sam_cam_fusion": args.sam_cam_fusion, "mga_pyramid_fusion": "add",
                   "parallelism": f"dp{world} (batch sharded, {args.scaling} scaling; all-reduce of {nparams} fp32 weight grads)",
                   "l2": "two rotating input/output sets per level (>= 0.73 GB touched per pair of steps) >> 126 MB L2",
                   "cuda_graph": main["graph"], "streams": main["streams"], "algorithmic_bytes_per_step": alg_bytes,
                   "launch_path": "one kernel per phase (MGA_FORCE_SPLIT)" if args.force_split else "cluster-per-sample kernels (cl_fwd / cl_bwd) + bwd_wgrad; shapes the cluster path does not take fall back to one kernel per phase",
                   "step_order": "forward of all levels (one stream per level), join, backward of all levels, join"},
        "gpu_launches": main["launches_per_step"] * args.steps,
        "launches_per_step": main["launches_per_step"],
        "roofline": roofline,
        "kernels": [{"kernel": k["kernel"], "level": k["level"], "ms": round(k["ms"], 5), "gbps": round(k.get("gbps", 0.0), 1),
                     "frac": round(k.get("frac", 0.0), 4)} for k in klist],
        "e2e": e2e,
        "clocks": clocks,
    }
    if variant is not None:
        line["variants"] = {other: {"ms_per_step": round(variant["ms"], 5), "value": round(world * alg_bytes / (variant["ms"] * 1e-3) / 1e9, 1),
                                    "images_per_sec": round(world * B / (variant["ms"] * 1e-3), 1),
                                    "note": "oracle: in-repo PyTorch composition; reference parity unpinned" if other == "add" else "reference-equivalent MaskCBAM"}}
    if variant_split is not None:
        line.setdefault("variants", {})["one_kernel_per_phase"] = {
            "ms_per_step": round(variant_split["ms"], 5), "value": round(world * alg_bytes / (variant_split["ms"] * 1e-3) / 1e9, 1),
            "launches_per_step": variant_split["launches_per_step"],
            "note": "MGA_FORCE_SPLIT: 12 kernels per level and direction; HBM traffic 10N instead of 5N"}
    if one_stream is not None:
        line.setdefault("variants", {})["one_stream"] = {
            "ms_per_step": round(one_stream["ms"], 5), "value": round(world * alg_bytes / (one_stream["ms"] * 1e-3) / 1e9, 1),
            "note": "the same step issued on ONE stream (fwd P3,P4,P5 then bwd P5,P4,P3), CUDA graph replay: what a training loop's module calls give"}
    if args.sam_cam_fusion == "add":
        line["config"]["note"] = "oracle: in-repo PyTorch composition; reference parity unpinned"
    solo = rank == 0 and world == 1 and not args.batch
    if solo and not args.no_cpu and args.workload == "cfg2":
        line["mask_pipeline"] = mask_pipeline(dev, B, 640, peak)
        line["inference_b1"] = inference_b1(dev, levels, dtype, os.cpu_count() or 1)
        try:
            line["spade_block"] = spade_block(dev, levels, B, dtype, peak)
        except Exception as e:  # pragma: no cover
            line["spade_block"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
        try:
            line["eca_block"] = eca_block(dev, levels, B, dtype, peak)
        except Exception as e:  # pragma: no cover
            line["eca_block"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    if solo and not args.no_workloads and args.workload == "cfg2":
        # the other BASELINE configs that fit one GPU: same step, same rules, fewer timed steps
        wl = {}
        for name in ("cfg3", "cfg5"):
            lv, Bw, dn, txt = WORKLOADS[name]
            dtw = DT[dn]
            ew = torch.empty((), dtype=dtw).element_size()
            r = measure(args, lib, dev, world, lv, Bw, dtw, flags_of(args.sam_cam_fusion), 10, 3)
            r["sets"] = None
            torch.cuda.empty_cache()
            ab = algorithmic_bytes(lv, Bw, ew)
            bw = annotate_kernels(r["kernels"], lv, Bw, ew, peak)
            wl[name] = {"workload": txt, "levels_CHW": lv, "batch": Bw, "dtype": dn, "ms_per_step": round(r["ms"], 5), "steps": 10, "warmup": 3,
                        "value": round(ab / (r["ms"] * 1e-3) / 1e9, 1), "unit": "GB/s", "images_per_sec": round(Bw / (r["ms"] * 1e-3), 1),
                        "step_frac": round(ab / (r["ms"] * 1e-3) / 1e9 / peak, 4), "launches_per_step": r["launches_per_step"],
                        "algorithmic_bytes_per_step": ab,
                        "dominant_kernel": None if bw is None else {"kernel": f"{bw['kernel']}[{bw['level']}]", "ms": round(bw["ms"], 5),
                                                                     "achieved": round(bw["gbps"], 1), "frac": round(bw["frac"], 4)},
                        "kernels": [{"kernel": k["kernel"], "level": k["level"], "ms": round(k["ms"], 5), "frac": round(k.get("frac", 0.0), 4)} for k in r["kernels"]]}
        bfp = None
        if peaks_path.exists():
            bfp = json.loads(peaks_path.read_text()).get("bf16_tflops_sustained")
        wl["cfg4"] = concat_workload(dev, peak, bfp)
        line["workloads"] = wl
    if args.workload == "cfg2" and not args.batch and not strong and not args.no_workloads and 256 % world == 0:
        # BASELINE configs[2] as it is stated: YOLOv8s bf16, GLOBAL batch 256 sharded over the ranks (strong scaling: 256 / N samples
        # per GPU), module-only fwd+bwd + the weight-gradient all-reduce; every rank takes part, rank 0 reports
        lv, Bg, dn, txt = WORKLOADS["cfg3"]
        dtw = DT[dn]
        ew = torch.empty((), dtype=dtw).element_size()
        Bs = Bg // world
        r = measure(args, lib, dev, world, lv, Bs, dtw, flags_of(args.sam_cam_fusion), 10, 3, instrument=False)
        r["sets"] = None
        torch.cuda.empty_cache()
        ab = algorithmic_bytes(lv, Bs, ew)
        line["strong_scaling_cfg3"] = {
            "workload": txt + f": global batch {Bg} sharded over {world} GPU(s)", "global_batch": Bg, "batch_per_gpu": Bs, "dtype": dn,
            "ms_per_step": round(r["ms"], 5), "steps": 10, "warmup": 3, "value": round(world * ab / (r["ms"] * 1e-3) / 1e9, 1), "unit": "GB/s",
            "images_per_sec": round(Bg / (r["ms"] * 1e-3), 1), "per_gpu_step_frac": round(ab / (r["ms"] * 1e-3) / 1e9 / peak, 4),
            "scaling": "strong", "timing": "CUDA events, max over ranks, all-reduce inside the timed region"}
    if solo and not args.no_cpu:
        line["gpu_eager_baseline"] = gpu_eager_baseline(dev, levels, B, dtype, alg_bytes, args.sam_cam_fusion)
        cb = 16 if args.workload == "cfg2" else 8
        sec, threads, kind = time_cpu(levels, cb, dtype, args.sam_cam_fusion, steps=3, warmup=1)
        cbytes = algorithmic_bytes(levels, cb, esize)
        line["cpu_baseline"] = {"value": round(cbytes / sec / 1e9, 3), "unit": "GB/s", "cores": threads, "kind": kind,
                                "sample": f"same workload at batch {cb} (3 timed steps after 1 warm-up, {sec*1e3:.1f} ms/step); "
                                          + ("the reference's MaskCBAM class (oracle/_ref) forward + torch autograd backward on CPU tensors" if kind == "reference"
                                             else "oracle/cbam_oracle.py forward + torch autograd backward"),
                                "images_per_sec": round(cb / sec, 1)}
    return line


def concat_workload(dev, peak, bf16_peak_tflops):
    """BASELINE configs[3]: YOLOv8m (reference-YAML widths 256/512/512), batch 128 bf16, sam_cam_fusion=concat + mga_pyramid_fusion=multiply,
    fwd+bwd through the public nn.Module.  These two modes have no reference source (parity unpinned, in-repo
    oracle).  Forward: CUDA gates op (s, a') + weight-folding pre-kernel + ONE tcgen05 kernel (virtual concat, both 1x1 convolutions,
    spatial gate, bias, pyramid fusion: csrc/cbam_concat.cuh).  Backward: three library GEMMs + one elementwise / reduction kernel +
    the gates backward.  `library_composition` times the same module as gates op + torch.cat + F.conv2d + autograd."""
    from mga_yolo_b200 import MaskGuidedCBAM

    levels, B, dtype = [(256, 80, 80), (512, 40, 40), (512, 20, 20)], 128, torch.bfloat16
    gen = torch.Generator(device=dev).manual_seed(5)
    work = []
    try:
        for (Cc, H, W) in levels:
            torch.manual_seed(Cc)
            mod = MaskGuidedCBAM(Cc, sam_cam_fusion="concat", mga_pyramid_fusion="multiply").to(dev)
            sets = [(torch.randn(B, Cc, H, W, generator=gen, device=dev).to(dtype), torch.randn(B, 1, H, W, generator=gen, device=dev),
                     torch.randn(B, Cc, H, W, generator=gen, device=dev).to(dtype)) for _ in range(2)]
            work.append((mod, sets))

        def step(i):
            for mod, sets in work:
                x, mk, g = sets[i & 1]
                xi = x.detach().requires_grad_(True)
                mi = mk.detach().requires_grad_(True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    out = mod([xi, mi])
                out.backward(g)
                mod.zero_grad(set_to_none=True)

        reps = 6

        def timed():
            for i in range(3):
                step(i)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                step(i)
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / reps

        def timed_graph():
            """The same two steps (input sets 0 / 1) captured in CUDA graphs and replayed: the device time of the step without the
            ~100 host-side calls per level, which bound the eager loop on boxes with a slow CPU (4.3 ms on one box, 6.3 ms on another)."""
            cap_s = torch.cuda.Stream(dev)
            torch.cuda.synchronize(dev)
            with torch.cuda.stream(cap_s):  # (the parameters' AccumulateGrad nodes must be created on the capture stream)
                for i in range(2):
                    step(i)
            torch.cuda.synchronize(dev)
            graphs = []
            for i in range(2):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=cap_s):
                    step(i)
                graphs.append(gph)
            with torch.cuda.stream(cap_s):
                for i in range(2):
                    graphs[i].replay()
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cap_s)
                for i in range(reps):
                    graphs[i & 1].replay()
                e1.record(cap_s)
                torch.cuda.synchronize(dev)
            del graphs
            return e0.elapsed_time(e1) / reps

        def best():
            eager = timed()
            try:
                return timed_graph(), eager, True
            except Exception as e:  # pragma: no cover
                print(f"[bench] cfg4 graph capture unavailable ({type(e).__name__}: {str(e)[:120]}); reporting the eager loop", file=sys.stderr)
                torch.cuda.synchronize(dev)
                return eager, eager, False

        ms, ms_eager, graphed = best()
        os.environ["MGA_CONCAT_LIBRARY"] = "1"
        try:
            ms_lib, ms_lib_eager, lib_graphed = best()
        finally:
            os.environ.pop("MGA_CONCAT_LIBRARY", None)
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    finally:
        work.clear()
        torch.cuda.empty_cache()
    ab = algorithmic_bytes(levels, B, 2)
    flops = sum(3 * 2.0 * (B * H * W) * (2 * Cc) * Cc for (Cc, H, W) in levels)  # forward GEMM + two backward GEMMs of the 2C->C 1x1 conv
    return {"workload": "BASELINE configs[3]: YOLOv8m (256/512/512 ch) batch 128 bf16, sam_cam_fusion=concat, mga_pyramid_fusion=multiply (module-only fwd+bwd)",
            "levels_CHW": levels, "batch": B, "dtype": "bfloat16", "ms_per_step": round(ms, 4), "steps": reps, "warmup": 3,
            "cuda_graph": graphed, "eager_ms_per_step": round(ms_eager, 4),
            "value": round(ab / (ms * 1e-3) / 1e9, 1), "unit": "GB/s", "step_frac": round(ab / (ms * 1e-3) / 1e9 / peak, 4),
            "images_per_sec": round(B / (ms * 1e-3), 1), "gemm_tflops": round(flops / (ms * 1e-3) / 1e12, 1),
            "gemm_frac_of_bf16_sustained": None if not bf16_peak_tflops else round(flops / (ms * 1e-3) / 1e12 / bf16_peak_tflops, 4),
            "path": "forward: gates op + fold + ONE tcgen05 kernel (TMA, TMEM accumulators, fused epilogue); backward: levels with C > 256 = ONE tcgen05 kernel (U, V in TMEM, closed form in the epilogue), C <= 256 = library GEMM + elementwise kernel; weight gradient = 2 per-sample library GEMMs (fp32) + 1 batch-reduce kernel; gates backward accumulates the concat dx (one autograd node)",
            "library_composition": {"ms_per_step": round(ms_lib, 4), "cuda_graph": lib_graphed, "eager_ms_per_step": round(ms_lib_eager, 4), "note": "same module as gates op + torch.cat + F.conv2d (cuDNN) + autograd (MGA_CONCAT_LIBRARY=1)"},
            "note": "oracle: in-repo PyTorch composition; reference parity unpinned"}


def gpu_eager_baseline(dev, levels, B, dtype, alg_bytes, sam_cam):
    """SURVEY.md section 8d's like-for-like comparator: the reference's MaskCBAM class (oracle/_ref) moved to the same B200 and run
    with eager PyTorch kernels (forward + autograd backward) on tensors of the same shapes.  ~200 ATen launches per level and direction."""
    if not reference_available() or sam_cam != "multiply":
        return {"unavailable": "oracle/_ref not built" if sam_cam == "multiply" else "the reference has no sam_cam_fusion=add"}
    from oracle import build_ref

    ref = build_ref.load()
    work = []
    gen = torch.Generator(device=dev).manual_seed(11)
    for (Cc, H, W) in levels:
        mod = ref.MaskCBAM(Cc)
        w1, b1, w2, b2, wsam, beta = make_params(Cc)
        mod.load_state_dict({"cam_mlp.0.weight": w1, "cam_mlp.0.bias": b1, "cam_mlp.2.weight": w2, "cam_mlp.2.bias": b2, "sam_conv.weight": wsam, "beta": beta})
        mod = mod.to(dev)
        x = torch.randn(B, Cc, H, W, generator=gen, device=dev).to(dtype)
        mk = torch.randn(B, 1, H, W, generator=gen, device=dev).to(dtype)
        g = torch.randn(B, Cc, H, W, generator=gen, device=dev).to(dtype)
        if dtype != torch.float32:
            mod = mod.to(dtype)  # module.bfloat16(): the reference's low-precision mode
        work.append((mod, x, mk, g))

    def step():
        for mod, x, mk, g in work:
            xi = x.detach().requires_grad_(True)
            mi = mk.detach().requires_grad_(True)
            mod([xi, mi]).backward(g)
            mod.zero_grad(set_to_none=True)

    try:
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
    except Exception as e:  # pragma: no cover  (e.g. out of memory at a large workload)
        return {"unavailable": repr(e)[:200]}
    finally:
        work.clear()
        torch.cuda.empty_cache()
    return {"ms_per_step": round(ms, 4), "value": round(alg_bytes / (ms * 1e-3) / 1e9, 1), "unit": "GB/s", "images_per_sec": round(B / (ms * 1e-3), 1),
            "steps": reps, "kind": "reference class on cuda:0, eager PyTorch (ATen/cuDNN kernels), same shapes and dtype, inputs resident in HBM"}


def bind_to_gpu_numa(local_rank):
    """Best effort: run this process on the CPUs of the NUMA node the GPU hangs off, BEFORE the pinned host buffers are
    allocated and first touched, so each rank's staging memory is local to its PCIe root (8 ranks x 0.7 GB/step otherwise share
    one socket's memory system).  Returns the node or None."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        devid = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(Path(path).read_text().strip())
        if node < 0:
            return None
        cpus = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        allowed = ids & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def pcie_probe(dev):
    """Raw pinned-host <-> device copy bandwidth of THIS box (256 MiB, copy streams): the e2e number is bounded by it, and the boxes differ
    (device->host was seen at 55 GB/s on most and ~20 GB/s on some)."""
    n = 256 << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a, d_b = torch.empty(n, dtype=torch.uint8, device=dev), torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(up, down, reps=4):
        def once():
            if up:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        once()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            once()
        for s_ in (s1, s2):
            torch.cuda.current_stream(dev).wait_stream(s_)
        e1.record()
        torch.cuda.synchronize(dev)
        return round((int(up) + int(down)) * n / (e0.elapsed_time(e1) / reps) / 1e6, 1)

    return {"h2d_GBps": run(True, False), "d2h_GBps": run(False, True), "both_GBps": run(True, True)}


def e2e_module(args, dev, local_rank, levels, B, dtype, world, alg_bytes):
    """Public API path with HOST buffers in and out: pinned host x / mask / grad_out -> H2D -> MaskGuidedCBAM forward ->
    autograd backward -> out, dx, dmask and the flat weight gradients -> D2H into pinned host memory.  Everything a caller of the
    reference's CPU path holds in host memory after a step is in host memory here too."""
    import torch.distributed as dist

    from mga_yolo_b200 import FlatGradReducer, MaskGuidedCBAM

    numa = bind_to_gpu_numa(local_rank)
    mods, host, devb, hres = [], [], [], []
    gen = torch.Generator().manual_seed(5)
    for (Cc, H, W) in levels:
        torch.manual_seed(Cc)
        m = MaskGuidedCBAM(Cc, sam_cam_fusion=args.sam_cam_fusion).to(dev)
        mods.append(m)
        hx = torch.randn(B, Cc, H, W, generator=gen).to(dtype).pin_memory()
        hm = torch.randn(B, 1, H, W, generator=gen).pin_memory()
        hg = torch.randn(B, Cc, H, W, generator=gen).to(dtype).pin_memory()
        host.append((hx, hm, hg))
        # two device input sets per level: the copies of step i + 1 run while step i computes and its results leave
        devb.append([(torch.empty_like(hx, device=dev), torch.empty_like(hm, device=dev), torch.empty_like(hg, device=dev)) for _ in range(2)])
        hres.append((torch.empty_like(hx).pin_memory(), torch.empty_like(hx).pin_memory(), torch.empty_like(hm).pin_memory()))  # out, dx, dmask
    reducer = FlatGradReducer([p for m in mods for p in m.parameters()])
    hgrad = torch.empty(reducer.numel, dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for trip in host for t in trip)
    d2h_full = hgrad.numel() * 4 + sum(t.numel() * t.element_size() for trip in hres for t in trip)

    # Input pipeline of a training loop: the host->device copies run on their own stream, level by level; the compute stream
    # waits for the copy of ITS level only, so the copies of the later levels (and of the next step) overlap the kernels; the
    # results leave on a third stream (PCIe is full duplex).  Each device buffer is rewritten only after its readers finished.
    main_s = torch.cuda.current_stream(dev)
    copy_s = torch.cuda.Stream(dev)
    back_s = torch.cuda.Stream(dev)
    ev_copied = [[torch.cuda.Event() for _ in levels] for _ in range(2)]
    ev_done = [[torch.cuda.Event() for _ in levels] for _ in range(2)]
    ev_back = [torch.cuda.Event() for _ in levels]
    for e in ev_done[0] + ev_done[1] + ev_back:
        e.record(main_s)
    keep = [None] * len(levels)  # results of the previous step stay alive until their D2H copy has been issued
    nstep = [0]

    def step(full):
        st_ = nstep[0] & 1
        nstep[0] += 1
        with torch.cuda.stream(copy_s):
            for li, ((hx, hm, hg), sets) in enumerate(zip(host, devb)):
                dx_, dm_, dg_ = sets[st_]
                copy_s.wait_event(ev_done[st_][li])  # the step before last has finished reading this set
                dx_.copy_(hx, non_blocking=True)
                dm_.copy_(hm, non_blocking=True)
                dg_.copy_(hg, non_blocking=True)
                ev_copied[st_][li].record(copy_s)
        for li, (m, sets) in enumerate(zip(mods, devb)):
            dx_, dm_, dg_ = sets[st_]
            main_s.wait_event(ev_copied[st_][li])
            xin = dx_.requires_grad_(True)
            min_ = dm_.requires_grad_(True)
            out = m([xin, min_])
            torch.autograd.backward(out, dg_, inputs=[xin, min_, *m.parameters()])
            gx, gm = xin.grad, min_.grad
            xin.grad = None
            min_.grad = None
            dx_.requires_grad_(False)
            dm_.requires_grad_(False)
            ev_done[st_][li].record(main_s)
            if full:
                ho, hdx, hdm = hres[li]
                back_s.wait_event(ev_done[st_][li])
                with torch.cuda.stream(back_s):
                    ho.copy_(out.detach(), non_blocking=True)
                    hdx.copy_(gx, non_blocking=True)
                    hdm.copy_(gm, non_blocking=True)
                    for t in (out, gx, gm):
                        t.record_stream(back_s)
                    ev_back[li].record(back_s)
            keep[li] = (out, gx, gm)
        if world > 1:
            reducer.all_reduce()
        hgrad.copy_(reducer.flat, non_blocking=True)

    host_ms = [0.0]

    def timed(full):
        steps = max(5, min(args.steps, 20))
        for _ in range(3):
            reducer.zero()
            step(full)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for _ in range(steps):
            reducer.zero()
            step(full)
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / steps  # CPU time to ENQUEUE a step (nothing waits inside the loop)
        for e in ev_back:
            main_s.wait_event(e)  # the closing event is after the last result has reached host memory
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            tmax = torch.tensor([ms], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax.item())
        return ms, steps

    ms_full, steps = timed(True)
    host_full = host_ms[0]
    eager_results = [tuple(t.clone() for t in trip) for trip in hres] + [hgrad.clone()]  # to check the graph replays against (deterministic kernels)
    ms_grads, _ = timed(False)

    # The same step captured in CUDA graphs (torch.cuda.graph around the public module calls and torch.autograd.backward), software
    # pipelined: graph p computes on device input set p and returns its results while it copies the NEXT step's inputs into set 1 - p.
    # One graph launch per step instead of ~100 host-side calls: the eager pipeline above is host bound on boxes with a slow CPU
    # (host_enqueue_ms_per_step), this one is bound by the PCIe link only.  Single-GPU runs only (the all-reduce stays eager).
    def graph_pipeline():
        cap_s = torch.cuda.Stream(dev)
        graphs, keepg = [], []
        torch.cuda.synchronize(dev)
        # the parameters' AccumulateGrad nodes remember the stream they were created on: drop every autograd graph of the eager steps
        # and run two steps on the capture stream first, so that the capture never touches the default stream
        for li in range(len(levels)):
            keep[li] = None
        with torch.cuda.stream(cap_s):
            for _ in range(2):
                reducer.zero()
                for m, sets in zip(mods, devb):
                    dx_, dm_, dg_ = sets[0]
                    xin, min_ = dx_.requires_grad_(True), dm_.requires_grad_(True)
                    out = m([xin, min_])
                    torch.autograd.backward(out, dg_, inputs=[xin, min_, *m.parameters()])
                    xin.grad = None
                    min_.grad = None
                    dx_.requires_grad_(False)
                    dm_.requires_grad_(False)
                    del out, xin, min_
        torch.cuda.synchronize(dev)
        for p in (0, 1):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=cap_s):
                cur = torch.cuda.current_stream(dev)
                copy_s.wait_stream(cur)
                with torch.cuda.stream(copy_s):  # next step's inputs
                    for (hx, hm, hg), sets in zip(host, devb):
                        for dst, src in zip(sets[1 - p], (hx, hm, hg)):
                            dst.copy_(src, non_blocking=True)
                reducer.zero()
                for li, (m, sets) in enumerate(zip(mods, devb)):
                    dx_, dm_, dg_ = sets[p]
                    xin, min_ = dx_.requires_grad_(True), dm_.requires_grad_(True)
                    out = m([xin, min_])
                    torch.autograd.backward(out, dg_, inputs=[xin, min_, *m.parameters()])
                    gx, gm = xin.grad, min_.grad
                    xin.grad = None
                    min_.grad = None
                    dx_.requires_grad_(False)
                    dm_.requires_grad_(False)
                    keepg.append((out, gx, gm))
                    back_s.wait_stream(cur)
                    with torch.cuda.stream(back_s):  # this level's results leave while the next level computes
                        ho, hdx, hdm = hres[li]
                        ho.copy_(out.detach(), non_blocking=True)
                        hdx.copy_(gx, non_blocking=True)
                        hdm.copy_(gm, non_blocking=True)
                hgrad.copy_(reducer.flat, non_blocking=True)
                cur.wait_stream(copy_s)
                cur.wait_stream(back_s)
            graphs.append(g)
        steps_g = max(5, min(args.steps, 20))
        with torch.cuda.stream(cap_s):
            for (hx, hm, hg), sets in zip(host, devb):  # inputs of the first step
                for dst, src in zip(sets[0], (hx, hm, hg)):
                    dst.copy_(src, non_blocking=True)
            k = 0
            for _ in range(4):  # warm-up replays (even count: the next replay is graph 0 again, its inputs prefetched by the last graph 1)
                graphs[k & 1].replay()
                k += 1
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cap_s)
            t_host = time.perf_counter()
            for _ in range(steps_g):
                graphs[k & 1].replay()
                k += 1
            host_g = (time.perf_counter() - t_host) * 1e3 / steps_g
            e1.record(cap_s)
            torch.cuda.synchronize(dev)
        same = all(torch.equal(a, b) for trip, ref in zip(hres, eager_results[:-1]) for a, b in zip(trip, ref)) and torch.equal(hgrad, eager_results[-1])
        if not same:
            raise RuntimeError("graph replays returned other results than the eager pipeline")
        return e0.elapsed_time(e1) / steps_g, host_g, keepg

    ms_graph = host_graph = None
    if world == 1:
        try:
            ms_graph, host_graph, _keep_alive = graph_pipeline()
        except Exception as e:  # pragma: no cover
            print(f"[bench] e2e graph pipeline unavailable ({type(e).__name__}: {e}); reporting the eager pipeline", file=sys.stderr)
            ms_graph = None
    eager = {"value": round(world * alg_bytes / (ms_full * 1e-3) / 1e9, 2), "ms_per_step": round(ms_full, 4), "host_enqueue_ms_per_step": round(host_full, 3),
             "note": "the same step issued call by call from Python (copy stream + compute stream + result stream, two device input sets)"}
    if ms_graph is not None:
        ms_full, host_full = ms_graph, host_graph
    try:
        probe = pcie_probe(dev)
    except Exception as e:  # pragma: no cover
        probe = {"error": str(e)}
    return {"value": round(world * alg_bytes / (ms_full * 1e-3) / 1e9, 2), "unit": "GB/s", "ms_per_step": round(ms_full, 4), "steps": steps,
            "pcie_probe": probe, "host_enqueue_ms_per_step": round(host_full, 3), "eager_pipeline": eager,
            "cuda_graph": ms_graph is not None,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_full, "images_per_sec": round(world * B / (ms_full * 1e-3), 1),
            "numa_node_of_pinned_buffers": numa,
            "api": "mga_yolo_b200.MaskGuidedCBAM forward + torch.autograd backward; pinned host x/mask/grad_out in (copy stream, two device input sets: the copies "
                   "of step i + 1 overlap the kernels and the result copies of step i); out, dx, dmask and the flat weight grads back to pinned host memory (second copy "
                   "stream, full duplex)" + ("; the step (copies included) is captured once with torch.cuda.graph around these public calls and replayed, one graph "
                                             "launch per step -- `eager_pipeline` is the same step issued call by call" if ms_graph is not None else ""),
            "result_stays_on_device": {"value": round(world * alg_bytes / (ms_grads * 1e-3) / 1e9, 2), "ms_per_step": round(ms_grads, 4),
                                       "d2h_bytes_per_step": hgrad.numel() * 4,
                                       "note": "same step when only the weight gradients return to the host (out feeds Detect and dx the neck's backward on the device, "
                                               "as inside the reference's training step)"}}


def mask_pipeline(dev, B, imgsz, peak):
    """Group 2 of the hot path: binary mask (B, imgsz, imgsz) uint8 -> the three pyramid masks (strides 8/16/32, default method =
    occupancy + 3x3 close, mga_yolo/data/dataset.py:95-103 + utils/mask_utils.py:64-141) in one kernel, next to the per-stride
    kernels and to the oracle port on the host (numpy restatement of the cv2 calls the reference makes per sample)."""
    import numpy as np

    from mga_yolo_b200 import MaskUtils
    from oracle import mask_oracle as mo

    rng = np.random.default_rng(0)
    nbuf = 6  # 6 x 26 MB of masks > 126 MB L2
    bufs = [torch.from_numpy((rng.random((B, imgsz, imgsz)) > 0.7).astype(np.uint8)).to(dev) for _ in range(nbuf)]
    alg = B * imgsz * imgsz + B * sum((imgsz // s) ** 2 for s in (8, 16, 32))

    def timed(fn, reps=30):
        for i in range(3):
            fn(bufs[i % nbuf])
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(bufs[i % nbuf])
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def timed_graph(fn, reps=30):
        """The same call on the 6 rotating batches captured in ONE CUDA graph (a dataloader processes batch after batch) and replayed: the
        kernels' time without the ~35 us of Python / dispatcher / allocator work per op call that bounds the eager loop at this size, and
        without one graph launch per 20-microsecond call."""
        for m in bufs:
            fn(m)
        torch.cuda.synchronize(dev)
        keep = []
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            for m in bufs:
                keep.append(fn(m))
        for _ in range(3):
            gph.replay()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            gph.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / (reps * nbuf)

    ms_one_eager = timed(lambda m: MaskUtils.masks_multi(m))
    ms_per = timed(lambda m: [MaskUtils.downsample_mask(m, s) for s in (8, 16, 32)])
    try:
        ms_one = timed_graph(lambda m: MaskUtils.masks_multi(m))
    except Exception as e:  # pragma: no cover
        print(f"[bench] mask pipeline graph capture failed ({e}); reporting the eager loop", file=sys.stderr)
        ms_one = ms_one_eager
    host = bufs[0][:8].cpu().numpy()
    kind, fn = "port", mo.downsample_mask
    if reference_available():
        try:  # the reference's own MaskUtils (three cv2 calls per sample in the dataloader workers, dataset.py:95-103)
            from oracle import build_ref

            kind, fn = "reference", build_ref.load().MaskUtils.downsample_mask
            fn(host[0], 8)
        except Exception:
            kind, fn = "port", mo.downsample_mask
    t0 = time.perf_counter()
    for b in range(host.shape[0]):
        for s in (8, 16, 32):
            fn(host[b], s)
    cpu_s = (time.perf_counter() - t0) / host.shape[0]
    return {"workload": f"{B} binary masks {imgsz}x{imgsz} uint8 -> strides 8/16/32, default method (block max + 3x3 close)",
            "one_pass_ms": round(ms_one, 5), "one_pass_eager_call_ms": round(ms_one_eager, 5), "per_stride_ms": round(ms_per, 5),
            "masks_per_sec": round(B / (ms_one * 1e-3), 1),
            "roofline": {"bound": "hbm", "achieved": round(alg / (ms_one * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(alg / (ms_one * 1e-3) / 1e9 / peak, 4), "alg_bytes_per_launch": alg,
                         "note": "two stages (mga_masks_multi_ws): one thread per 8x8 block over the whole batch reads the masks, then one CTA per (8-row band, "
                                 "image, stride) derives the maps + 3x3 close; 6 rotating batches captured in one CUDA graph, replayed (the eager op call is "
                                 "bounded by ~35 us of host work)"},
            "cpu_baseline": {"masks_per_sec": round(1.0 / cpu_s, 1), "cores": 1, "kind": kind,
                             "sample": "8 masks x 3 strides, " + ("the reference's MaskUtils.downsample_mask (cv2) from oracle/_ref" if kind == "reference"
                                                                  else "oracle/mask_oracle.py (numpy restatement of cv2.resize / morphologyEx)")}}


def spade_block(dev, levels, B, dtype, peak):
    """SURVEY 8f-4 neighbour: the feature side of MaskSPADE (instance statistics + gamma * xhat + beta, and its closed-form backward) through
    torch.ops.mga.spade_fwd / spade_bwd at the same pyramid shapes; gamma / beta stand for the outputs of the block's mask branch.
    Algorithmic bytes per level: forward x, gamma, beta in + y out = 4 N e; backward x, g, gamma in + dx, d gamma out = 5 N e."""
    e = torch.empty((), dtype=dtype).element_size()
    sets = []
    for (C, H, W) in levels:
        gen = torch.Generator(device=dev).manual_seed(C)
        sets.append(tuple(torch.randn(B, C, H, W, device=dev, dtype=dtype, generator=gen) for _ in range(4)))  # x, gamma, beta, g
    alg = sum(9 * B * C * H * W * e for (C, H, W) in levels)

    def step():
        for x, gm, bt, g in sets:
            _, stats = torch.ops.mga.spade_fwd(x, gm, bt, 1e-6)
            torch.ops.mga.spade_bwd(g, x, gm, stats, True)

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    reps = 10
    per = []
    for li, (x, gm, bt, g) in enumerate(sets):  # per-kernel durations: back-to-back launches of one kernel (P3 touches 0.4-0.5 GB per launch: nothing stays in L2)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        _, stats = torch.ops.mga.spade_fwd(x, gm, bt, 1e-6)
        torch.cuda.synchronize(dev)
        ev[0].record()
        for _ in range(reps):
            torch.ops.mga.spade_fwd(x, gm, bt, 1e-6)
        ev[1].record()
        for _ in range(reps):
            torch.ops.mga.spade_bwd(g, x, gm, stats, True)
        ev[2].record()
        torch.cuda.synchronize(dev)
        tf, tb = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        n = B * levels[li][0] * levels[li][1] * levels[li][2] * e
        per.append({"level": f"P{3 + li}", "fwd_ms": round(tf / reps, 5), "bwd_ms": round(tb / reps, 5),
                    "fwd_frac": round(4 * n / (tf / reps * 1e-3) / 1e9 / peak, 4), "bwd_frac": round(5 * n / (tb / reps * 1e-3) / 1e9 / peak, 4)})
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    # comparator: the same math as the reference block runs it (masked_spade.py:126,143): F.instance_norm + gamma * xhat + beta under torch autograd
    lib_ms = None
    try:
        import torch.nn.functional as F

        leaves = [tuple(t.clone().requires_grad_(True) for t in (x, gm, bt)) + (g,) for x, gm, bt, g in sets]

        def lib_step():
            for x, gm, bt, g in leaves:
                y = gm * F.instance_norm(x, eps=1e-6) + bt
                torch.autograd.backward(y, g, inputs=[x, gm, bt])
                x.grad = gm.grad = bt.grad = None

        for _ in range(2):
            lib_step()
        torch.cuda.synchronize(dev)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(5):
            lib_step()
        l1.record()
        torch.cuda.synchronize(dev)
        lib_ms = l0.elapsed_time(l1) / 5
        del leaves
    except Exception as ex:  # pragma: no cover
        print(f"[bench] spade library comparator failed: {ex}", file=sys.stderr)
    return {"library_composition_ms_per_step": None if lib_ms is None else round(lib_ms, 4),
            "workload": f"MaskSPADE feature side (mga_spade_forward / mga_spade_backward), levels {levels}, batch {B}, {str(dtype).split('.')[-1]}",
            "ms_per_step": round(ms, 5), "value": round(alg / (ms * 1e-3) / 1e9, 1), "unit": "GB/s", "algorithmic_bytes_per_step": alg,
            "roofline": {"bound": "hbm", "achieved": round(alg / (ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s", "frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4)},
            "kernels": per, "launches_per_step": 2 * len(levels),
            "note": "eager op calls on one stream (6 launches per step); the mask branch that produces gamma / beta is library convolution work and not timed here"}


def eca_block(dev, levels, B, dtype, peak):
    """SURVEY 8f-4 neighbour: MaskECA forward + backward (mga_eca_forward / mga_eca_backward through the nn.Module) at the same pyramid
    shapes.  Algorithmic bytes per level: forward (2 N + B S) e, backward (3 N + 2 B S) e  (x, mask in, out; x, g in, dx, dmask out)."""
    from mga_yolo_b200 import MaskECA

    e = torch.empty((), dtype=dtype).element_size()
    items = []
    for (C, H, W) in levels:
        gen = torch.Generator(device=dev).manual_seed(C)
        mod = MaskECA(C).to(dev)
        x = torch.randn(B, C, H, W, device=dev, dtype=dtype, generator=gen).requires_grad_(True)
        m = torch.randn(B, 1, H, W, device=dev, generator=gen).requires_grad_(True)
        g = torch.randn(B, C, H, W, device=dev, dtype=dtype, generator=gen)
        items.append((mod, x, m, g))
    alg = sum((5 * B * C * H * W) * e + 3 * B * H * W * 4 for (C, H, W) in levels)

    def step():
        for mod, x, m, g in items:
            out = mod([x, m])
            torch.autograd.backward(out, g, inputs=[x, m, *mod.parameters()])
            x.grad = m.grad = None

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    gph = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(dev)
    with torch.cuda.stream(cap):
        step()
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(gph, stream=cap):
            step()
        for _ in range(2):
            gph.replay()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(cap)
        for _ in range(reps):
            gph.replay()
        e1.record(cap)
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return {"workload": f"MaskECA block (nn.Module -> mga_eca_forward / mga_eca_backward), levels {levels}, batch {B}, {str(dtype).split('.')[-1]}",
            "ms_per_step": round(ms, 5), "value": round(alg / (ms * 1e-3) / 1e9, 1), "unit": "GB/s", "algorithmic_bytes_per_step": alg,
            "roofline": {"bound": "hbm", "achieved": round(alg / (ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s", "frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4)},
            "note": "module calls + torch.autograd.backward on one stream, captured in a CUDA graph and replayed; 8 kernels per level"}


def inference_b1(dev, levels, dtype, threads):
    """BASELINE configs[0] at module level: the three MaskCBAM calls of one YOLOv8n inference (imgsz 640, batch 1), eval + no_grad,
    through the public nn.Module (MGA_NO_SAVE forward), next to the oracle port on the host."""
    from mga_yolo_b200 import MaskGuidedCBAM
    from oracle import cbam_oracle as co

    mods, xs, ms = [], [], []
    gen = torch.Generator().manual_seed(7)
    for (Cc, H, W) in levels:
        torch.manual_seed(Cc)
        mods.append(MaskGuidedCBAM(Cc).to(dev).eval())
        xs.append(torch.randn(1, Cc, H, W, generator=gen).to(dtype))
        ms.append(torch.randn(1, 1, H, W, generator=gen))
    dx = [t.to(dev) for t in xs]
    dm = [t.to(dev) for t in ms]

    def fwd():
        with torch.no_grad():
            return [m([x, k]) for m, x, k in zip(mods, dx, dm)]

    for _ in range(5):
        fwd()
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fwd()
    reps = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_graph = e0.elapsed_time(e1) / reps
    t0 = time.perf_counter()
    for _ in range(50):
        fwd()
    torch.cuda.synchronize(dev)
    ms_eager = (time.perf_counter() - t0) / 50 * 1e3
    torch.set_num_threads(threads)
    ps = [co.CbamParams(*make_params(Cc)) for (Cc, _, _) in levels]
    with torch.no_grad():
        for _ in range(2):
            [co.cbam_forward(x.float(), k, p)[0] for x, k, p in zip(xs, ms, ps)]
        t0 = time.perf_counter()
        for _ in range(10):
            [co.cbam_forward(x.float(), k, p)[0] for x, k, p in zip(xs, ms, ps)]
        cpu_ms = (time.perf_counter() - t0) / 10 * 1e3
    return {"workload": "BASELINE configs[0] at module level: the three MaskCBAM forwards of one YOLOv8n inference, imgsz 640, batch 1, eval + no_grad",
            "gpu_ms_cuda_graph": round(ms_graph, 5), "gpu_ms_eager_module_calls": round(ms_eager, 4),
            "cpu_baseline": {"ms": round(cpu_ms, 3), "cores": threads, "kind": "port", "sample": "10 forwards, oracle/cbam_oracle.py"}}


def reference_arm(args, rank):
    """CPU arm: the reference's own MaskCBAM class (oracle/_ref; the oracle port when that is absent or for the `add` mode,
    which the reference does not have) on the host cores with all host threads.  cfg2 runs the FULL batch 64 and honours
    --steps / --warmup (same config, same steps as the GPU arm); the larger workloads run a bounded batch-8 sample."""
    if rank != 0:
        return None
    levels, B, dtname, desc_txt = WORKLOADS[args.workload]
    dtype = DT[dtname]
    esize = torch.empty((), dtype=dtype).element_size()
    cb = args.batch or (B if args.workload == "cfg2" else 8)
    steps, warmup = args.steps, args.warmup
    if args.workload != "cfg2":
        steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 3))
    sec, threads, kind = time_cpu(levels, cb, dtype, args.sam_cam_fusion, steps=steps, warmup=warmup)
    cbytes = algorithmic_bytes(levels, cb, esize)
    val = round(cbytes / sec / 1e9, 3)
    sample = f"batch {cb} of the workload per step ({steps} timed steps after {warmup} warm-up), all {threads} host threads"
    return {
        "impl": "reference", "metric": "mga_cbam_fwd_bwd_algorithmic_GBps", "value": val, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": {"float32": "f32", "bfloat16": "bf16", "float16": "f16"}[dtname],
        "data": "synthetic (same distributions and parameter values as the GPU arm)",
        "images_per_sec": round(cb / sec, 1),
        "config": {"workload": desc_txt, "levels_CHW": levels, "batch_per_gpu": cb, "global_batch": cb, "sam_cam_fusion": args.sam_cam_fusion,
                   "mga_pyramid_fusion": "add"},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("the reference's unmodified MaskCBAM class (mga_yolo/nn/modules/masked_cbam.py, laid out under oracle/_ref by oracle/build_ref.py) "
                 "forward + torch autograd backward on CPU tensors") if kind == "reference" else
                "oracle/cbam_oracle.py (pinned against reference-generated goldens) with torch autograd backward",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--sam-cam-fusion", default="multiply", choices=["multiply", "add"],
                    help="multiply = the reference's MaskCBAM (parity pinned); add = BASELINE's build-side variant")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--one-stream", action="store_true", help="run the three levels back to back on one stream")
    ap.add_argument("--no-variant", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the host-buffer end-to-end measurement")
    ap.add_argument("--no-workloads", action="store_true", help="skip the cfg3 / cfg5 block of the default (cfg2) run")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank runs the workload's full batch; strong: the workload's batch is the GLOBAL batch, sharded over the ranks")
    ap.add_argument("--force-split", action="store_true", help="never use the cluster-resident fused kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout carries exactly ONE line (the JSON record): everything libraries print while we run (NCCL's version banner,
    # Ultralytics-style loggers ...) goes to stderr -- file descriptor 1 is pointed at stderr until the record is ready.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        line = reference_arm(args, rank)
        if line is not None:
            emit(line)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. The B200 path has no CPU fallback; use --impl reference for the CPU arm.")
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = gpu_arm(args, rank, world, local_rank)
    if rank == 0:
        emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
