#!/usr/bin/env python
"""Per CUDA source line totals (stall samples, warp instructions) from
`ncu --page source --csv --print-source cuda,sass`.  usage: python tools/ncu_lines.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = {}
cur_file = ""
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("Warp Stall Sampling (All Samples)"); i_ex = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= i_ex or not r[0].isdigit():
        continue
    if r[2] != "-":          # SASS rows carry an address; source rows carry '-'
        continue
    key = (cur_file, int(r[0]))
    s, ex = int(r[i_s] or 0), int(r[i_ex] or 0)
    a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
    a[0] += s; a[1] += ex
ts = sum(v[0] for v in agg.values()) or 1
te = sum(v[1] for v in agg.values()) or 1
print("total samples %d, warp instructions %d" % (ts, te))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% stall %5.1f%% inst  %s:%d  %s" % (100.0 * v[0] / ts, 100.0 * v[1] / te, k[0], k[1], v[2]))
