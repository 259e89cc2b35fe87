#!/usr/bin/env python
"""Per-source-line shares of warp-stall samples and executed instructions from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > cs.csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
agg = {}
fname = ""
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ismp = hdr.index("# Samples")
        iex = hdr.index("Instructions Executed")
        continue
    if hdr is None or not r[0] or not r[0].isdigit():
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [r[1], 0, 0])
    a[1] += int(r[ismp]) if r[ismp].isdigit() else 0
    a[2] += int(r[iex]) if r[iex].isdigit() else 0
tots = sum(v[1] for v in agg.values()) or 1
toti = sum(v[2] for v in agg.values()) or 1
print("samples", tots, "instructions", toti)
for key, v in sorted(agg.items()):
    if v[1] / tots > thr or v[2] / toti > thr:
        print("%-22s %4d %5.1f%% smp %5.1f%% inst | %s" % (key[0][:22], key[1], 100 * v[1] / tots, 100 * v[2] / toti, v[0].strip()[:100]))
