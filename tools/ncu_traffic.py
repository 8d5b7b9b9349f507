"""From an ncu CSV with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch (one training
step, tools/gpu_step_once.py): per-kernel-family time and DRAM traffic, and the implicit-GEMM aggregate that
bench.py reports as roofline.traffic.
    python tools/ncu_traffic.py gpurun_out/step_launches.csv profiles/r01_conv_traffic.json > profiles/r01_step_kernels.txt"""
import csv
import json
import re
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0,
        "ms": 1e3, "msecond": 1e3}

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, mi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit", "ID"))
launch = defaultdict(dict)
names = {}
for row in r:
    if len(row) <= vi:
        continue
    names[row[ii]] = re.sub(r"^void ", "", re.sub(r"\(.*", "", row[ki]))
    launch[row[ii]][row[mi]] = float(row[vi].replace(",", "")) * UNIT.get(row[ui], 1.0)
agg = defaultdict(lambda: [0, 0.0, 0.0])
for lid, m in launch.items():
    a = agg[names[lid]]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
tot_us = sum(a[1] for a in agg.values())
print("one training step: %d launches, %.2f ms summed device time (ncu: serialised, cold caches: compare SHARES)" % (
    sum(a[0] for a in agg.values()), tot_us / 1e3))
print("%7s %10s %6s %10s %8s  %s" % ("share", "us", "n", "DRAM MB", "GB/s", "kernel"))
for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%% %10.1f %6d %10.1f %8.0f  %s" % (100 * us / tot_us, us, n, by / 1e6, by / us / 1e3 if us else 0, k))
gemm = [(n, us, by) for k, (n, us, by) in agg.items() if re.search(r"conv_halo2?_kernel|conv_fwd_kernel|wgrad", k)
        and "unpack" not in k]
n = sum(g[0] for g in gemm)
out = {"source": sys.argv[1], "launches": n, "dram_bytes_total": sum(g[2] for g in gemm),
       "dram_bytes_per_launch": sum(g[2] for g in gemm) / max(n, 1), "device_us_total": sum(g[1] for g in gemm),
       "share_of_step": sum(g[1] for g in gemm) / tot_us}
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
print("implicit-GEMM launches: %d, DRAM traffic %.2f GB per step (%.1f MB per launch), share of the step's device time "
      "%.1f%%" % (n, out["dram_bytes_total"] / 1e9, out["dram_bytes_per_launch"] / 1e6, 100 * out["share_of_step"]))
