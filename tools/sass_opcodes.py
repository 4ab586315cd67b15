#!/usr/bin/env python
"""Static SASS opcode counts per kernel of the shipped libpqc_b200.so (cuobjdump -sass) -> the table of
profiles/r2_sass_opcodes.txt.  python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyramaterised_b200", "libpqc_b200.so")
COLS = ["DFMA", "DMUL", "DADD", "DMMA", "LDGSTS", "UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "SHFL",
        "BAR", "IMAD", "LOP3", "BRA", "LDL", "STL"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
            counts[cur]["total"] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    for n, d in zip(names, dem):
        demangle[n] = re.sub(r"\(.*", "", d).replace("void ", "")
    print("# SASS opcode counts per kernel of the shipped libpqc_b200.so (cuobjdump -sass; static counts; "
          "tools/sass_opcodes.py)")
    print("# DMMA = FP64 tensor core (mma.sync.m8n8k4.f64), LDGSTS = cp.async, UBLKCP = cp.async.bulk (TMA "
          "bulk copy), SYNCS = mbarrier")
    print("kernel | total | " + " | ".join(COLS))
    for n in sorted(names, key=lambda k: -counts[k]["total"]):
        c = counts[n]
        print(" | ".join([demangle[n], str(c["total"])] + [str(c[k]) for k in COLS]))


if __name__ == "__main__":
    main()
