"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, data = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr:
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    a = agg.setdefault(d["Kernel Name"], [0, 0.0])
    a[0] += 1
    a[1] += float(d["Metric Value"].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {a[0]:8d} {a[1]/1e3:10.1f} {a[1]/a[0]/1e3:9.1f} {a[1]/tot*100:6.1f}%")
