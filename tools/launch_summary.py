"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    if unit == "ns":
        v /= 1e3
    elif unit == "ms":
        v *= 1e3
    agg[name][0] += 1
    agg[name][1] += v
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot/1e3:.2f} ms of kernel time (cold-cache, serialised: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} n={v[0]:5d}  total={v[1]/1e3:9.2f} ms  avg={v[1]/v[0]:9.1f} us  share={v[1]/tot*100:5.1f}%")
