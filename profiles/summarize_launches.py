"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    try:
        v = float(row[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if row[ui] == "ns" else (v * 1e3 if row[ui] == "ms" else v)
    k = row[ki].split("(")[0]
    acc[k][0] += 1
    acc[k][1] += v
tot = sum(v[1] for v in acc.values())
print("kernel,launches,total_us,share")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{v[0]},{v[1]:.1f},{v[1] / tot:.4f}")
