#!/usr/bin/env python
"""Summaries of ncu captures for profiles/ (run here, on the CPU box, on files brought back in gpurun_out/).

  tools/ncu_summary.py launches LAUNCHES.csv           per-kernel totals/shares of a gpu__time_duration launch list
  tools/ncu_summary.py full REPORT.ncu-rep             per-launch key metrics of an `ncu --set full` report
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("%-28s %5s %12s %10s %7s" % ("kernel", "n", "total_ms", "avg_us", "share"))
    for k, a in agg.items():
        print("%-28s %5d %12.3f %10.1f %6.1f%%" % (k, a[0], a[1] / 1e6, a[1] / a[0] / 1e3, 100 * a[1] / tot))
    print("%-28s %5s %12.3f" % ("total", "", tot / 1e6))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(k), n) for k, n in KEYS if k in hdr]
    kn = hdr.index("Kernel Name")
    print("%-22s " % "kernel" + " ".join("%9s" % n for _, n in cols))
    print("%-22s " % "" + " ".join("%9s" % units[i][:9] for i, _ in cols))
    for d in data:
        name = d[kn].split("(")[0].replace("void ", "")[:22]
        vals = []
        for i, _ in cols:
            try:
                v = float(d[i].replace(",", ""))
                vals.append("%9.3f" % v if v < 1e4 else "%9.3g" % v)
            except ValueError:
                vals.append("%9s" % d[i][:9])
        print("%-22s " % name + " ".join(vals))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
