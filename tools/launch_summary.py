"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel."""
import collections, csv, sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
        continue
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    unit = row["Metric Unit"]
    v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    name = row["Kernel Name"][:100]
    agg[name][0] += 1
    agg[name][1] += v
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot / 1e3:.2f} ms of kernel time")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{v[1] / 1e3:9.2f} ms {v[0]:6d}x {100 * v[1] / tot:5.1f}%  {k}")
