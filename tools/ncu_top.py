"""Print the hottest SASS instructions (warp-stall samples) of an .ncu-rep source page CSV.
    ncu -i rep --page source --csv > x.csv ; python tools/ncu_top.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
si = hdr.index("Warp Stall Sampling (All Samples)")
ex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((float(r[si]), k, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for s, k, r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + float(r[i] or 0)
print("stall reasons:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for s, k, r in sorted(data, key=lambda t: -t[0])[:n]:
    top = sorted(((float(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%7.0f %5.1f%%  #%-5d exec %-9s %-60s %s" % (s, 100 * s / tot, k, r[ex], r[1].strip()[:60],
                                                      ",".join("%s:%d" % (h[6:], v) for v, h in top if v)))
