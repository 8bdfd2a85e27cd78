"""Host-side cost of one e2e step of bench.py (enqueue time without synchronisation) + cProfile of 20 steps."""
import cProfile, pstats, sys, time, types
import torch
sys.path.insert(0, ".")
import bench

args = types.SimpleNamespace(sam_cam_fusion="multiply", steps=20)
dev = torch.device("cuda:0")
levels, B, dtname, _ = bench.WORKLOADS["cfg2"]
# re-use e2e_module's internals by monkeypatching its timed(): we only want the step closure -> copy the function source is overkill; time the whole call instead
bench.e2e_module(args, dev, 0, levels, B, bench.DT[dtname], 1, 923955200)  # imports, allocator, caches
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
out = bench.e2e_module(args, dev, 0, levels, B, bench.DT[dtname], 1, 923955200)
pr.disable()
print("e2e", out["value"], out["ms_per_step"], "wall", round(time.perf_counter() - t0, 2), "s")
pstats.Stats(pr).sort_stats("tottime").print_stats(32)
