#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page source --csv`: top SASS instructions by warp-stall samples with their dominant
stall reasons.  Usage: ncu -i rep --page source --csv | python tools/ncu_top_stalls.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = list(csv.reader(sys.stdin))
hdr = next(r for r in rows if r and r[0] == "Address")
start = rows.index(hdr) + 1
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data, tot = [], 0
for r in rows[start:]:
    try:
        s = int(r[ci["# Samples"]])
    except (ValueError, IndexError):
        continue
    tot += s
    reasons = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    data.append((s, r[ci["Address"]][-5:], r[ci["Source"]], [(c, v) for v, c in reasons if v]))
data.sort(key=lambda x: -x[0])
print(f"# total samples {tot}")
for s, a, src, rs in data[:n]:
    print(f"{s:7d} {100.0 * s / max(1, tot):5.1f}% {a} {src[:90]:90s} {rs}")
