"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: python tools/summarize_launches.py gpurun_out/launches.csv [--all]"""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = [(r["Kernel Name"].split("(")[0].split("::")[-1], float(r["Metric Value"].replace(",", "")) / 1e3, r["Grid Size"])
        for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
tot = sum(x[1] for x in rows)
agg = collections.OrderedDict()
for k, v, g in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
print(f"total {tot / 1e3:.3f} ms over {len(rows)} launches")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} n={n:4d} {v:10.1f} us {100 * v / tot:5.1f}%")
if "--all" in sys.argv:
    for i, (k, v, g) in enumerate(rows):
        print(f"{i:4d} {k:34s} {v:9.1f} us grid {g}")
