"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    unit = row["Metric Unit"]
    t = t / 1000.0 if unit == "ns" else t * 1000.0 if unit == "ms" else t
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1000:.2f} ms over {sum(v[0] for v in agg.values())} launches ({path})")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1] / 1000:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d} avg {v[1] / v[0]:8.1f} us  {k}")
