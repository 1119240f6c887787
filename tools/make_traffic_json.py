"""profiles/r02_traffic.json from an ncu capture of ONE step of a bench workload:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --csv ... tools/profile_step.py
    python tools/make_traffic_json.py gpurun_out/<tag>_traffic.csv c2 profiles/r02_traffic.json
Per kernel family (the families bench.py times live): launches per step, DRAM bytes per launch, L2 bytes, ncu time.
bench.py prints `roofline.traffic` from this file only if the launch count matches the live one."""
import collections
import csv
import json
import sys

FAMILY = [("conv_halo", "conv_halo"), ("conv_umma", "conv_tap"), ("conv_stem", "stem_conv"), ("stem_grad", "stem_grad"),
          ("pool3s1_fwd", "pool_fwd"), ("maxpool_fwd", "pool_fwd"), ("pool3s1_bwd", "pool_bwd"), ("pool_s2_bwd", "pool_bwd"),
          ("maxpool_bwd", "pool_bwd"), ("apply", "apply"), ("delta_update", "delta_update"), ("head_", "head_loss"),
          ("loss_kernel", "head_loss"), ("stem_bias", "other")]


def main():
    path, config, out = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, vi, ii, ui = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[ii], {"kernel": r[ki]})
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        if "byte" in u:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1e-3)
        d[r[mi]] = v
    fam = collections.OrderedDict()
    for d in per.values():
        name = next((f for pat, f in FAMILY if pat in d["kernel"]), "other")
        a = fam.setdefault(name, {"launches": 0, "dram_bytes_per_step": 0.0, "l2_bytes_per_step": 0.0, "ncu_us_per_step": 0.0})
        a["launches"] += 1
        a["dram_bytes_per_step"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["l2_bytes_per_step"] += d.get("lts__t_bytes.sum", 0.0)
        a["ncu_us_per_step"] += d.get("gpu__time_duration.sum", 0.0)
    for a in fam.values():
        a["dram_bytes_per_launch"] = a["dram_bytes_per_step"] / a["launches"]
    json.dump({"config": config, "families": fam,
               "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum over every launch of one "
                         f"step of bench config {config} (tools/profile_step.py; {path})"}, open(out, "w"), indent=1)
    for k, a in fam.items():
        print(f"{k:14s} launches {a['launches']:3d}  dram {a['dram_bytes_per_step'] / 1e6:9.1f} MB/step  "
              f"{a['dram_bytes_per_launch'] / 1e6:8.1f} MB/launch  L2 {a['l2_bytes_per_step'] / 1e6:9.1f} MB  ncu {a['ncu_us_per_step']:8.1f} us")


if __name__ == "__main__":
    main()
