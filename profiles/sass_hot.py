#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel launch from `ncu -i X.ncu-rep --page source --csv` output.
usage: sass_hot.py file.csv [section_index] [top_n]"""
import csv, sys
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
s = secs[k]; e = secs[k + 1] if k + 1 < len(secs) else len(rows)
H = rows[s + 1]
body = [r for r in rows[s + 2:e] if len(r) == len(H)]
iS = H.index("# Samples"); iI = H.index("Instructions Executed"); iSrc = H.index("Source")
stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in body)
print("kernel:", rows[s][1][:90], "| sections:", len(secs), "| instrs:", len(body), "| samples:", tot,
      "| warp-instr executed:", sum(int(r[iI]) for r in body))
agg = {}
for r in body:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[i])
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
order = sorted(range(len(body)), key=lambda i: -int(body[i][iS]))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[c]), h[6:]) for c, h in stall_cols if int(r[c])), reverse=True)[:3]
    print(f"{i:5d} {int(r[iS]):6d} {100*int(r[iS])/tot:5.1f}% x{r[iI]:>8s}  {r[iSrc].strip()[:70]:70s} {st}")
