"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.
    python tools/ncu_launch_summary.py gpurun_out/launches_r01.csv > profiles/r01_launches_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
ui = hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for row in r:
    if len(row) <= vi or "gpu__time_duration" not in row[mi]:
        continue
    v = float(row[vi].replace(",", ""))
    unit = row[ui]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*", "", row[ki])
    name = re.sub(r"^void ", "", name)
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print("launches %d, summed device time %.2f ms (ncu serialises launches and runs them cold-cache: compare SHARES)" % (
    sum(v[0] for v in agg.values()), tot / 1000.0))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%%  %9.2f us  %5d x  %s" % (100.0 * us / tot, us, n, k))
