#!/usr/bin/env python
"""Join an `ncu --page source --csv` dump (SASS + executed counts + stall samples) with the line
table of the SAME build (nvdisasm -g of the kernel's cubin, extracted here from the in-tree .so)
and report executed instructions / samples per SOURCE line and per opcode class.
  python tools/ncu_line_summary.py source.csv KERNEL_MANGLED [top]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(mangled):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "pyramaterised_b200", "libpqc_b200.so")],
                   cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for cub in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
        key = f".text.{mangled}:"
        if key not in txt:
            continue
        lines = txt.split("\n")
        start = next(i for i, l in enumerate(lines) if key in l)
        table, cur = {}, ("?", 0)
        for l in lines[start + 1:]:
            if l.startswith("//---------------------"):
                break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
            if m:
                table[int(m.group(1), 16)] = (cur, m.group(2))
        return table
    raise SystemExit("kernel not found in the library")


def main():
    src, mangled = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    table = line_table(mangled)
    csv.field_size_limit(1 << 30)
    rows = list(csv.reader(open(src)))
    k = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name")
    hdr = rows[k + 1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[k + 2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) >= len(hdr):
            data.append(r)
    base = int(data[0][idx["Address"]], 16)
    ex, sm, cls = collections.Counter(), collections.Counter(), collections.Counter()
    per_line_cls = collections.defaultdict(collections.Counter)
    for r in data:
        off = int(r[idx["Address"]], 16) - base
        line, op = table.get(off, (("?", 0), "?"))
        n = int(r[idx["Instructions Executed"]])
        ex[line] += n
        sm[line] += int(r[idx["# Samples"]])
        c = "fp64" if op[:4] in ("DFMA", "DMUL", "DADD") else ("mov" if op.startswith("IMAD.MOV") or op == "MOV"
                                                                else ("lds/sts" if op[:3] in ("LDS", "STS") else
                                                                      ("branch" if op[:3] in ("BRA", "BSS", "BSY") else "other")))
        cls[c] += n
        per_line_cls[line][c] += n
    tot, tots = sum(ex.values()), sum(sm.values())
    print("executed warp-instructions", tot, " classes:",
          ", ".join(f"{c} {100 * v / tot:.1f}%" for c, v in cls.most_common()))
    print("line: executed% samples% [fp64 mov lds branch other]")
    for line, v in ex.most_common(top):
        c = per_line_cls[line]
        print(f"  {line[0]}:{line[1]:<5d} {100 * v / tot:5.1f}% {100 * sm[line] / max(1, tots):5.1f}%  "
              f"[{c['fp64']} {c['mov']} {c['lds/sts']} {c['branch']} {c['other']}]")


if __name__ == "__main__":
    main()
