"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the first launches."""
import collections
import csv
import sys

path = sys.argv[1]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(d["Kernel Name"], float(d["Metric Value"]) / 1000.0, d.get("Grid Size")) for d in csv.DictReader(lines)]
agg = collections.OrderedDict()
for n, t, g in rows:
    a = agg.setdefault(n[:70], [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} n={a[0]:5d} total={a[1]:10.1f} us avg={a[1] / a[0]:8.1f} us {100 * a[1] / tot:5.1f}%")
for n, t, g in rows[:first]:
    print(f"  {n[:56]:58s} {t:8.1f} us grid={g}")
