"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): total time and share per kernel."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void\s+|nepb::", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v_us
    a[2] = max(a[2], v_us)
tot = sum(a[1] for a in agg.values())
print("%-60s %8s %12s %8s %10s %10s" % ("kernel", "launches", "total us", "share", "avg us", "max us"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %8d %12.1f %7.1f%% %10.2f %10.1f" % (k[:60], a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2]))
print("%-60s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
