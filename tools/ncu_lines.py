"""Stall samples per CUDA source line of an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py <report.ncu-rep> [top lines]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None
per = collections.OrderedDict()
H = None
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No" or (len(r) > 3 and "Warp Stall Sampling (All Samples)" in r):
        H = r; continue
    if H is None:
        continue
    # cuda line rows: first col = line number, second = source; sass rows follow with Address
    if r[0].isdigit() and len(r) >= 2 and (len(r) < 4 or not r[2].startswith("0x")):
        cur = (fname, int(r[0]), r[1].strip()[:110])
        per.setdefault(cur, collections.Counter())
        # aggregated metrics may be on this row as well
        if len(r) == len(H):
            for k in ("Warp Stall Sampling (All Samples)", "Instructions Executed"):
                v = r[H.index(k)]
                if v.isdigit():
                    per[cur][k] += int(v)
            for i, h in enumerate(H):
                if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit():
                    per[cur][h] += int(r[i])
tot = sum(c["Warp Stall Sampling (All Samples)"] for c in per.values())
toti = sum(c["Instructions Executed"] for c in per.values())
print(f"total samples {tot}, warp instructions {toti/1e6:.2f} M")
key = "Instructions Executed" if "--inst" in sys.argv else "Warp Stall Sampling (All Samples)"
items = sorted(per.items(), key=lambda kv: -kv[1][key])[:top]
for (f, ln, src), c in sorted(items, key=lambda kv: (kv[0][0], kv[0][1])):
    s = c["Warp Stall Sampling (All Samples)"]
    reasons = sorted(((v, k[6:]) for k, v in c.items() if k.startswith("stall_")), reverse=True)[:3]
    print(f"{f}:{ln:4d} {100*s/max(tot,1):5.1f}% inst {c['Instructions Executed']/1e3:7.1f}k  {' '.join(f'{k}:{v}' for v,k in reasons):40s} | {src}")
