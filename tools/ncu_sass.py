"""Summarise the SASS page of an ncu report: per-opcode executed warp instructions and the hottest straight-line blocks.
usage: python tools/ncu_sass.py <report.ncu-rep> [top blocks]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
si, ei, wi = H.index("Source"), H.index("Instructions Executed"), H.index("Warp Stall Sampling (All Samples)")
ins = [(r[si].strip(), int(r[ei]), int(r[wi])) for r in rows[hdr + 1:] if len(r) > wi and r[ei].isdigit()]
tot = sum(e for _, e, _ in ins)
samp = sum(s for _, _, s in ins)
print(f"{len(ins)} SASS instructions, {tot/1e6:.2f} M warp instructions executed, {samp} stall samples")
ops = collections.Counter()
for s, e, _ in ins:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    ops[op.split(".")[0]] += e
print("by opcode:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in ops.most_common(22)))
# blocks = maximal runs with the same execution count (+-0)
blocks, cur = [], None
for i, (s, e, w) in enumerate(ins):
    if cur is None or e != cur[2]:
        if cur:
            blocks.append(cur)
        cur = [i, i, e, 0, 0]
    cur[1] = i
    cur[3] += e
    cur[4] += w
blocks.append(cur)
blocks.sort(key=lambda b: -b[3])
for b in blocks[:top]:
    i0, i1, e, te, tw = b
    oc = collections.Counter((ins[i][0].split()[0] if not ins[i][0].startswith("@") else ins[i][0].split()[1]).split(".")[0] for i in range(i0, i1 + 1))
    print(f"  sass[{i0}:{i1}] n={i1-i0+1:4d} x {e:8d} = {te/1e6:6.2f} M ({100*te/tot:4.1f}%)  stalls {100*tw/max(samp,1):4.1f}%  " + " ".join(f"{k}:{v}" for k, v in oc.most_common(8)))
