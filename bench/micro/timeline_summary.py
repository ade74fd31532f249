"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` log of bench/micro/hex_timeline.py: kernels of the second solve."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/hex_tl.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1000.0, r["Grid Size"])
        for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
second = rows[len(rows) // 2:]
agg = collections.OrderedDict()
for n, v, g in second:
    a = agg.setdefault(n.split("(")[0][-48:], [0, 0.0])
    a[0] += 1
    a[1] += v
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:50s} n={a[0]:3d} total={a[1]:8.1f} us  avg={a[1] / a[0]:7.1f}")
print("total us", round(sum(a[1] for a in agg.values()), 1), "launches", len(second))
