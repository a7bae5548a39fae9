#!/usr/bin/env python3
"""Top stall sites of one kernel from `ncu -i rep --page source --csv` output:  python profiles/stalls.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
def f(x):
    try: return int(float(x))
    except ValueError: return 0
body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Address": break   # second (source-level) table
    if len(r) == len(h): body.append(r)
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
print("samples", sum(f(r[si]) for r in body), "instructions", len(body), "warp-instr executed", sum(f(r[ie]) for r in body))
agg = {h[i]: sum(f(r[i]) for r in body) for i in stall}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for n, r in enumerate(body): r.append(n)
for r in sorted(body, key=lambda r: -f(r[si]))[:topn]:
    st = sorted(((h[i], f(r[i])) for i in stall if f(r[i]) > 0), key=lambda kv: -kv[1])[:3]
    print(str(r[-1]).rjust(5), r[si].rjust(6), r[ie].rjust(9), r[src][:56].ljust(56), st)
