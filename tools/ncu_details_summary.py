"""Condense `ncu --page details --csv` exports (tools/ncu_families.sh) into one table:
    python tools/ncu_details_summary.py gpurun_out/r02_ncu_*.details.csv > profiles/r02_ncu_kernels.md"""
import csv
import sys

WANT = ["Duration", "DRAM Throughput", "Memory Throughput", "Compute (SM) Throughput", "Registers Per Thread",
        "Dynamic Shared Memory Per Block", "Achieved Occupancy", "L2 Hit Rate", "Executed Ipc Active", "Grid Size"]
print("| capture | kernel | " + " | ".join(WANT) + " |")
print("|---|---|" + "---|" * len(WANT))
for path in sys.argv[1:]:
    rows = list(csv.DictReader(open(path)))
    if not rows:
        continue
    vals = {}
    for r in rows:
        n = r.get("Metric Name")
        if n in WANT and n not in vals:
            vals[n] = (r.get("Metric Value", ""), r.get("Metric Unit", ""))
    kern = rows[0].get("Kernel Name", "")[:70]
    name = path.split("r02_ncu_")[-1].replace(".details.csv", "")
    print("| %s | `%s` | " % (name, kern) + " | ".join("%s %s" % vals.get(w, ("", "")) for w in WANT) + " |")
