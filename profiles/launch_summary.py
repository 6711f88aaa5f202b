"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel:
   python profiles/launch_summary.py gpurun_out/launches_vN.csv [first_id last_id]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi, idi = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
lo, up = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 10**9)
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi or not (lo <= int(r[idi]) <= up):
        continue
    a = agg.setdefault(r[ki][:120], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{'us':>10s} {'n':>5s} {'share':>6s}  kernel")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e3:10.1f} {c:5d} {100 * t / tot:5.1f}%  {n}")
print(f"{tot / 1e3:10.1f} total")
