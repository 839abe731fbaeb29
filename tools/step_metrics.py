#!/usr/bin/env python
"""Per-kernel LSU / DRAM / issue utilisation of one training step from an `ncu --metrics ... --csv` launch log
(the sweep that found the LSU-bound BatchNorm and classifier kernels in r01h).

    ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,\\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,\\
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file X.csv \\
        python bench.py --steps 1 --warmup 3 --ncu-range
    python tools/step_metrics.py X.csv > profiles/rNN_step_metrics.txt
"""
import collections
import csv
import sys

T, L, D, I = ("gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active")


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[ii], r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    for (_, k), m in per.items():
        name = k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
        a, t = agg[name], m[T]
        a[0] += 1
        a[1] += t
        a[2] += t * m[L]
        a[3] += t * m[D]
        a[4] += t * m[I]
        a[5] += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
    total = sum(a[1] for a in agg.values())
    print(f"# {sum(a[0] for a in agg.values())} launches, {total / 1e6:.2f} ms summed device time (ncu: serialised, cold cache); "
          "time-weighted averages per kernel")
    print(f"{'kernel':70s} {'n':>4s} {'ms':>8s} {'share':>6s} {'lsu%':>6s} {'dram%':>6s} {'issue%':>6s} {'dram GB':>8s}")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
        print(f"{n:70s} {a[0]:4d} {a[1] / 1e6:8.3f} {100 * a[1] / total:5.1f}% {a[2] / a[1]:6.1f} {a[3] / a[1]:6.1f} {a[4] / a[1]:6.1f} {a[5] / 1e9:8.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
