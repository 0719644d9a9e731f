"""python tools/summarize_launches.py launches.csv  -> per-kernel totals of an ncu gpu__time_duration launch list.
Kernels of libfslic_b200.so (k_*) are one iterate(); anything else is torch synthesising the input images."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else v * (1000 if row["Metric Unit"] == "ms" else 1)
    a = agg[row["Kernel Name"].split("(")[0][:44]]
    a[0] += 1
    a[1] += v
ours = {k: v for k, v in agg.items() if k.replace("void ", "").startswith("k_")}
other = {k: v for k, v in agg.items() if k not in ours}
tot = sum(v[1] for v in ours.values())
print("%-46s %5s %10s %9s %6s" % ("kernel (one iterate)", "n", "total_us", "avg_us", "share"))
for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print("%-46s %5d %10.1f %9.1f %5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
print("%-46s %5d %10.1f" % ("TOTAL (serialised, cold cache)", sum(v[0] for v in ours.values()), tot))
if other:
    print("\nnot part of a step (torch kernels that synthesise the input batch): %d launches, %.1f us"
          % (sum(v[0] for v in other.values()), sum(v[1] for v in other.values())))
