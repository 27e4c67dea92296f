"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", "")) * scale.get(row["Metric Unit"], 1.0)
    except (ValueError, KeyError):
        continue
    agg.setdefault(row["Kernel Name"].split("(")[0], []).append(v)
total = sum(sum(v) for v in agg.values())
print(f"{'kernel':45s} {'n':>5s} {'total ms':>10s} {'mean ms':>10s} {'max ms':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:45s} {len(v):5d} {sum(v):10.3f} {sum(v)/len(v):10.4f} {max(v):10.3f} {sum(v)/total:7.3f}")
print(f"{'TOTAL':45s} {sum(len(v) for v in agg.values()):5d} {total:10.3f}")
