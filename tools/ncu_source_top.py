#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from `ncu --page source --csv`.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name K --launch-count 1 > src.csv; python tools/ncu_source_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Address')
hdr = rows[h]
i_src, i_s, i_ex, i_thr = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed')
data = []
for n, r in enumerate(rows[h + 1:]):
    if len(r) > i_thr and r[0].startswith('0x'):
        data.append((int(r[i_s] or 0), int(r[i_ex] or 0), r[i_thr], r[i_src].strip(), n))
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "total warp-inst", sum(d[1] for d in data), "sass lines", len(data))
for d in sorted(data, key=lambda d: -d[0])[:top]:
    print("%6d %5.1f%% ex=%9d thr=%5s  #%4d %s" % (d[0], 100 * d[0] / tot, d[1], d[2], d[4], d[3][:100]))
