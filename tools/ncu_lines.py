#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export: share of stall samples and of executed
instructions for every line above a threshold.   python tools/ncu_lines.py gpurun_out/x_cs.csv [min_pct]"""
import collections
import csv
import sys

csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cur, hdr, last = None, None, None
agg = collections.OrderedDict()
stall_cols = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        si, ie = hdr.index('# Samples'), hdr.index('Instructions Executed')
        stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
        continue
    if r[0] == 'Function Name' or hdr is None:
        continue
    try:
        s, n = int(r[si]), int(r[ie])
    except (ValueError, IndexError):
        continue
    if r[0]:
        last = (cur, r[0])
        agg.setdefault(last, [0, 0, r[1], collections.Counter()])
    a = agg[last]
    a[0] += s
    a[1] += n
    for i, h in stall_cols.items():
        try:
            a[3][h] += int(r[i])
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()) or 1
tin = sum(a[1] for a in agg.values()) or 1
print("samples", tot, "warp-instructions", tin)
for k, a in agg.items():
    if 100 * a[0] / tot >= thr or 100 * a[1] / tin >= 2 * thr:
        top = ",".join(f"{h[6:]}:{100 * v / max(1, a[0]):.0f}" for h, v in a[3].most_common(2))
        print(f"{k[0]}:{k[1]:>4} s={100 * a[0] / tot:5.1f}% i={100 * a[1] / tin:5.1f}% [{top}] {a[2].strip()[:100]}")
