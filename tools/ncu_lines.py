#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).

usage: tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top_n]
Prints, per CUDA source line, the warp-stall samples and the warp-level instructions executed,
sorted by samples — the view used to decide what to fix in a kernel."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg, cur_file, hdr = {}, None, None
    tot_s = tot_i = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and len(r) > 7:
            try:
                s, i = int(r[6] or 0), int(r[7] or 0)
            except ValueError:
                continue
            key = (cur_file, int(r[0]))
            if key in agg:
                agg[key][0] += s
                agg[key][1] += i
            else:
                agg[key] = [s, i, r[1].strip()]
            tot_s += s
            tot_i += i
    print("kernel %s: %d samples, %d warp instructions (all matching launches summed)" % (kern, tot_s, tot_i))
    lines = [(v[0], v[1], k[0], k[1], v[2]) for k, v in agg.items()]
    for s, i, f, ln, src in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%-4d %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), f, ln, src[:110]))


if __name__ == "__main__":
    main()
