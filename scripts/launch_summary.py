"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python scripts/launch_summary.py launches.csv [skip_first_n] > summary.txt
Kernel names are shortened to their function name; conv / wgrad launches are further split by grid size."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1], newline="") as f:
    for r in csv.reader(f):
        if len(r) > 14 and r[12] == "gpu__time_duration.sum":
            rows.append(r)
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4]
    m = re.search(r"([A-Za-z_0-9:]+)\s*(<|\()", name)
    short = m.group(1) if m else name[:60]
    if "at::" in name or "elementwise" in name:
        short = "torch:" + (re.search(r"(\w+_kernel\w*|\w+Kernel\w*)", name).group(1) if re.search(r"(\w+_kernel\w*|\w+Kernel\w*)", name) else short)
    agg[short][0] += 1
    agg[short][1] += float(r[14].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print("%d launches, %.3f ms total" % (len(rows), tot / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%5d  %10.1f us  %5.1f %%  avg %8.1f us" % (k[:60], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
