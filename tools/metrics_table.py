"""Per-launch table from tools/gpu_ncu_metrics.sh output: python tools/metrics_table.py gpurun_out/metrics.csv"""
import collections, csv, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = collections.OrderedDict()
for r in csv.DictReader(lines):
    k = int(r["ID"])
    d = rows.setdefault(k, {"name": r["Kernel Name"].split("(")[0].split("::")[-1][:26], "grid": r["Grid Size"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = r["Metric Unit"]
    n = r["Metric Name"]
    if n == "gpu__time_duration.sum":
        d["us"] = v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    elif n.startswith("sm__pipe_tensor"):
        d["tc%"] = v
    elif n.startswith("lts__throughput"):
        d["l2%"] = v
    elif n.startswith("dram__throughput"):
        d["dram%"] = v
    elif n == "dram__bytes_read.sum":
        d["rdMB"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
    elif n == "dram__bytes_write.sum":
        d["wrMB"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
    elif n.startswith("l1tex__m_xbar2l1tex"):
        d["l2rdMB"] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
print(f"{'#':>3} {'kernel':26s} {'us':>8} {'tc%':>5} {'l2%':>5} {'dram%':>5} {'rdMB':>7} {'wrMB':>7} {'l2rdMB':>8}")
for i, (k, d) in enumerate(rows.items()):
    print(f"{i:3d} {d['name']:26s} {d.get('us',0):8.1f} {d.get('tc%',0):5.1f} {d.get('l2%',0):5.1f} {d.get('dram%',0):5.1f} "
          f"{d.get('rdMB',0):7.1f} {d.get('wrMB',0):7.1f} {d.get('l2rdMB',0):8.1f}")
