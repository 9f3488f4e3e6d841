#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: the SASS instructions holding most stall samples.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 | python tools/ncu_hot.py [min_pct]"""
import csv
import sys

rows = [r for r in csv.reader(sys.stdin)]
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
ix = {k: i for i, k in enumerate(hdr)}
minp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.4
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("kernel", rows[0][1][:90] if rows[0] else "", "samples", tot, "instructions", len(data))
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[ix[k]] or 0) for r in data) for k in keys}
print("stall totals:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
cum = 0
for n, r in enumerate(data):
    s = int(r[ix["# Samples"]])
    cum += s
    if s > tot * minp / 100:
        st = sorted(((k[6:], int(r[ix[k]] or 0)) for k in keys), key=lambda kv: -kv[1])[:2]
        print(f"{n:5d} {100 * s / tot:5.2f}% cum {100 * cum / tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {st}")
