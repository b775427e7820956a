#!/usr/bin/env python
"""Summarises ncu outputs brought back from the GPU box into small text files for profiles/.

    python scripts/ncu_summary.py launches gpurun_out/launches.csv > profiles/<name>_launches.txt
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/<name>_full.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r["Kernel Name"].split("(")[0][:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum` (cold-cache, serialised:")
    print("# compare SHARES, not absolutes).  total = %.1f us over %d launches" % (total / 1e3, sum(a[0] for a in agg.values())))
    print("%-92s %6s %12s %8s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-92s %6d %12.1f %8.1f %6.1f%%" % (name, n, ns / 1e3, ns / n / 1e3, 100 * ns / total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print("== kernel:", r[idx["Kernel Name"]][:100], " grid", r[idx.get("Grid Size", 0)], " block", r[idx.get("Block Size", 0)])
        for k in KEYS:
            if k in idx:
                print("   %-70s %16s %s" % (k, r[idx[k]], units[idx[k]]))
        if "dram__bytes_read.sum" in idx:
            def val(k):
                v = float(r[idx[k]].replace(",", ""))
                u = units[idx[k]]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            print("   %-70s %16.1f MB" % ("traffic = dram read + write", (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / 1e6))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
