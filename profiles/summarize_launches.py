"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python profiles/summarize_launches.py launches.csv [first_id last_id]"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    rows.append((int(r["ID"]), r["Kernel Name"], us))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = [r for r in rows if lo <= r[0] <= hi]
agg = defaultdict(lambda: [0, 0.0])
for _, name, us in rows:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
print(f"launches {len(rows)}  total {tot:.1f} us  (ids {rows[0][0]}..{rows[-1][0]})")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100 * us / tot:5.1f}%  n={n:4d}  avg {us / n:8.2f} us  {name[:110]}")
