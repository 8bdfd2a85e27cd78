"""bench.py's cfg4 block alone. usage: python tools/cfg4_only.py"""
import json, sys, torch
sys.path.insert(0, ".")
import bench
print(json.dumps({k: v for k, v in bench.concat_workload(torch.device("cuda:0"), 6453.4, 1396.0).items() if k not in ("path", "note", "workload")}))
