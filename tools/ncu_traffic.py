#!/usr/bin/env python
"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every stage's kernel from `ncu --set full`
reports -> the JSON bench.py reads for roofline.traffic.

usage: tools/ncu_traffic.py OUT.json NOTE REPORT.ncu-rep [REPORT2.ncu-rep ...]"""
import csv
import json
import subprocess
import sys

STAGE = {"k_fast_strips": "fast", "k_quadtree": "quadtree", "k_blur": "blur", "k_orient_describe": "orient_describe",
         "k_cape_sums": "cells", "k_cape_fit": "fit", "k_cape_edges": "fit", "k_cape_grid": "grid", "k_cape_refine_plan": "refine", "k_cape_paint": "refine",
         "k_cape_refine_border": "refine",
         "k_pyr_stream": "pyramid", "k_pyr_level0": "pyramid"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    out, note, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    stages = {}
    for rep in reps:
        rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
        hdr, units = rows[0], rows[1]
        kn, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        ia = hdr.index("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")
        ii = hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in hdr else -1
        iw_ = hdr.index("smsp__inst_executed.sum") if "smsp__inst_executed.sum" in hdr else -1
        for d in rows[2:]:
            name = d[kn].split("(")[0].replace("void ", "").replace("drfe::", "")
            base = name.split("<")[0]
            if base not in STAGE:
                continue
            b = float(d[ir].replace(",", "")) * UNIT[units[ir]] + float(d[iw].replace(",", "")) * UNIT[units[iw]]
            t = float(d[it].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[it], 1.0)
            s = stages.setdefault(STAGE[base], {"kernel": name, "dram_bytes": 0.0, "ncu_time_us": 0.0, "launches": 0})
            s["dram_bytes"] += b
            s["ncu_time_us"] += t
            s["launches"] += 1
            s["alu_pipe_pct"] = max(s.get("alu_pipe_pct", 0.0), float(d[ia].replace(",", "")))
            if ii >= 0:
                s["issue_pct"] = max(s.get("issue_pct", 0.0), float(d[ii].replace(",", "")))
            if iw_ >= 0:
                s["warp_inst"] = s.get("warp_inst", 0.0) + float(d[iw_].replace(",", ""))
    json.dump({"source": note, "stages": stages}, open(out, "w"), indent=1)
    print(json.dumps(stages, indent=1))


if __name__ == "__main__":
    main()
