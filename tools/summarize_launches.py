"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
    k = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
    tot[k][0] += 1
    tot[k][1] += v
s = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'n':>6s} {'total_us':>12s} {'avg_us':>9s} {'share':>6s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {v[0]:6d} {v[1]:12.1f} {v[1] / v[0]:9.2f} {v[1] / s:6.3f}")
print(f"{'TOTAL':70s} {sum(v[0] for v in tot.values()):6d} {s:12.1f}")
