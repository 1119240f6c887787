"""Print a compact per-launch table from an .ncu-rep: python tools/ncu_table.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_executed.sum", "inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active", "umma%"),
        ("sm__inst_executed_pipe_uniform.sum", "uni"),
        ("launch__registers_per_thread", "regs")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
units = rows[1]
print(" | ".join(n for _, n in idx))
for r in rows[2:]:
    vals = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0][-28:]
        else:
            try:
                f = float(v.replace(",", ""))
                u = units[i]
                if n == "us":
                    f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                if n in ("rdMB", "wrMB"):
                    f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
                v = f"{f:.1f}" if n != "inst" else f"{f/1e6:.1f}M"
            except ValueError:
                pass
        vals.append(v)
    print(" | ".join(vals))
