#!/usr/bin/env python
"""Summarises ncu outputs brought back in gpurun_out/ into the small text files committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python profiles/summarize.py full gpurun_out/x.ncu-rep         > profiles/rNN_x_full.txt
"""
import csv
import collections
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    return re.sub(r"\(.*", "", name)[:90]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(path, errors="replace")) if r and r[0] == "ID")
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)  # -> us
        k = short(r[ki])
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += v
        total += v
    print(f"# {len(rows)} launches, {total / 1e3:.3f} ms summed device time (ncu: cold-cache, serialised -> compare SHARES)")
    print(f"{'kernel':90s} {'n':>6s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:90s} {n:6d} {us / 1e3:10.3f} {us / n:10.1f} {100 * us / total:6.2f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", short(r[hdr.index("Kernel Name")]))
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:75s} {r[i]:>16s} {units[i]}")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        print(f"  traffic = dram read + write = {float(r[rd].replace(',', '')) + float(r[wr].replace(',', '')):.4f} {units[rd]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
