#!/usr/bin/env python
"""Warp-sample split of an ncu source page by SASS region:  ncu -i X.ncu-rep --page source --csv --print-source sass > f.csv;
python tools/ncu_roles.py f.csv [lo:hi ...]   (regions as instruction index ranges; without them prints 100-instruction bins
and the SYNCS / UTC* / LDTM instructions that mark the role boundaries)"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
S, E = ix["# Samples"], ix["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", sum(int(r[S]) for r in data), "instructions", sum(int(r[E]) for r in data))
if len(sys.argv) == 2:
    for a in range(0, len(data), 100):
        print(a, sum(int(r[S]) for r in data[a:a + 100]))
    for i, r in enumerate(data):
        src = r[ix["Source"]]
        if any(k in src for k in ("SYNCS.PHASECHK", "UTCBAR", "LDTM", "UCGABAR")) or int(r[S]) > 100:
            print(i, r[S], src[:90])
for rg in sys.argv[2:]:
    a, b = (int(x) for x in rg.split(":"))
    c = Counter()
    for r in data[a:b]:
        for s in stalls:
            c[s[6:]] += int(r[ix[s]])
    n = sum(int(r[S]) for r in data[a:b])
    print(rg, "samples", n, "instr", sum(int(r[E]) for r in data[a:b]), [(k, v) for k, v in c.most_common(7)])
