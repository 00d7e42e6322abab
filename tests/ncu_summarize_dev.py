"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel calls / avg us."""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
for k, (n, t) in agg.items():
    print("%-70s calls %4d  avg %10.1f us  total %10.1f us" % (k[:70], n, t / n, t))
