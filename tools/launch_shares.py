#!/usr/bin/env python
"""profiles/<tag>_launches.csv (ncu --metrics gpu__time_duration.sum of a short bench run) -> per-kernel totals and the share of
k_scan in one device-resident step.  usage: python tools/launch_shares.py profiles/r02_launches.csv > profiles/r02_launch_shares.txt"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
L = [(re.sub(r"\(.*", "", r[4]).replace("void ", ""), int(r[6]), int(r[8].strip("()").split(",")[0]), float(r[14]) / 1e6) for r in rows]
print("# %s = ncu --metrics gpu__time_duration.sum --clock-control none -c 400 on" % sys.argv[1])
print("#   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extra   (tools/final_measure.sh; times are cold-cache and serialised)")
print("# The run holds the pre-flight file scan, 3 + 2 + 2 device-resident steps (k_scan + k_xa each), the tuple-path comparison and the BGZF scans.\n")
print("## all launches of the run")
tot = sum(x[3] for x in L)
agg = {}
for k, _, _, ms in L:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s n=%4d  %10.3f ms  %5.1f%%" % (k, n, ms, 100 * ms / tot))
print("\n## one device-resident step of the bench (`value`): k_scan over the whole resident stream, then k_xa on its (empty) queue")
big = max((x[3] for x in L if x[0].startswith("k_scan")), default=0)
for i, x in enumerate(L):
    if x[0].startswith("k_scan") and x[3] > 0.5 * big and i + 1 < len(L) and L[i + 1][0] == "k_xa":
        print("k_scan %.3f ms + k_xa %.3f ms  -> k_scan is %.1f %% of the step's kernel time" % (x[3], L[i + 1][3], 100 * x[3] / (x[3] + L[i + 1][3])))
print("\n## one BGZF scan (`e2e`): the k_inflate launches of a 3.4 GB file (serialised under the profiler; they overlap in the real run)")
inf = [x for x in L if x[0].startswith("k_inflate") or x[0].startswith("k_lz")]
print("k_inflate / k_lz launches: %d, %.1f ms in all" % (len(inf), sum(x[3] for x in inf)))
