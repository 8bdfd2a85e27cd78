"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count / mean us."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
seq = []
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v
    seq.append((r[ki].split("(")[0].replace("void ", "").replace("mga::", ""), v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for n, v in seq[skip:]:
    a = agg.setdefault(n, [])
    a.append(v)
tot = sum(sum(v) for n, v in agg.items() if "at::" not in n)
for n, v in agg.items():
    if "at::" in n:
        continue
    print(f"{len(v):4d} x {sum(v)/len(v):8.2f} us  ({100*sum(v)/tot:5.1f}%)  min {min(v):7.2f}  {n[:70]}")
print(f"total {tot:.1f} us over {sum(len(v) for n, v in agg.items() if 'at::' not in n)} launches")
