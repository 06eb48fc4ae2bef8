#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of libdrfe.so (cuobjdump -sass), for profiles/rNN_sass_opcodes.txt: the evidence
behind the pipe arguments of DESIGN.md (VIMNMX / VABSDIFF4 / SHF / LOP3 = the half-rate ALU pipe, IMAD / FFMA = the FMA pipe,
I2F / F2I / MUFU = the XU pipe, UBLKCP / LDGSTS = bulk and asynchronous copies).

usage: tools/sass_hist.py [LIB.so] [KERNEL_REGEX] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dr-slam_b200", "libdrfe.so")
    want = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
    kernels, cur, k = collections.OrderedDict(), None, 0
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(re.sub(r"\(.*", "", name[k]).replace("void ", "").replace("drfe::", ""), collections.Counter())
            k += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op, mods = m.group(1), m.group(2)
            key = op + (mods if op in ("VIMNMX", "VIMNMX3", "VABSDIFF4", "IDP", "LDGSTS", "UBLKCP", "I2F", "F2I", "I2FP", "F2F", "MUFU") else "")
            cur[key] += 1
    for kn, c in kernels.items():
        if want and not want.search(kn):
            continue
        tot = sum(c.values())
        print("%s: %d instructions" % (kn, tot))
        print("   " + "  ".join("%s %d" % (o, n) for o, n in c.most_common(28)))


if __name__ == "__main__":
    main()
