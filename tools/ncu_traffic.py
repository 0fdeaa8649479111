"""Summarise an ncu launch list with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per kernel name:
time share, DRAM bytes and the DRAM rate each kernel type runs at.  Usage: ncu_traffic.py launches.csv [first] [count]"""
import csv, sys, collections, re
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
agg = collections.OrderedDict()
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
for row in csv.DictReader(lines):
    i = int(row["ID"])
    if i < first or i >= first + count:
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void\s+|nepb::", "", name)
    a = agg.setdefault(name, {"ids": set(), "t": 0.0, "r": 0.0, "w": 0.0})
    a["ids"].add(i)
    v = float(row["Metric Value"].replace(",", "")) * scale.get(row["Metric Unit"], 1.0)
    m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        a["t"] += v
    elif m == "dram__bytes_read.sum":
        a["r"] += v
    elif m == "dram__bytes_write.sum":
        a["w"] += v
tt = sum(a["t"] for a in agg.values())
tb = sum(a["r"] + a["w"] for a in agg.values())
print("%-44s %7s %10s %6s %10s %10s %6s %8s" % ("kernel", "launch", "time us", "share", "read MB", "write MB", "share", "GB/s"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    b = a["r"] + a["w"]
    print("%-44s %7d %10.1f %5.1f%% %10.1f %10.1f %5.1f%% %8.0f" % (k[:44], len(a["ids"]), a["t"], 100 * a["t"] / tt, a["r"] / 1e6, a["w"] / 1e6,
                                                               100 * b / max(tb, 1), b / 1e3 / max(a["t"], 1e-9)))
print("%-44s %7d %10.1f        %10.1f %10.1f        %8.0f" % ("TOTAL", sum(len(a["ids"]) for a in agg.values()), tt, sum(a["r"] for a in agg.values()) / 1e6,
                                                         sum(a["w"] for a in agg.values()) / 1e6, tb / 1e3 / tt))
