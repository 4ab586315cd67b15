#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall reasons, samples per opcode and the
hottest SASS lines of the first kernel in the file.  python tools/ncu_source_summary.py f.csv [N]"""
import collections
import csv
import sys


def main():
    f = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    csv.field_size_limit(1 << 30)
    rows = list(csv.reader(open(f)))
    k = 0
    while k < len(rows):
        if rows[k] and rows[k][0] == "Kernel Name":
            name = rows[k][1]
            hdr = rows[k + 1]
            idx = {h: i for i, h in enumerate(hdr)}
            data = []
            k += 2
            while k < len(rows) and rows[k] and rows[k][0] != "Kernel Name":
                if len(rows[k]) >= len(hdr):
                    data.append(rows[k])
                k += 1
            report(name, hdr, idx, data, top)
            break
        k += 1


def report(name, hdr, idx, data, top):
    S = idx["# Samples"]
    tot = sum(int(r[S]) for r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(int(r[idx[s]]) for r in data) for s in stalls}
    print(name, "SASS lines", len(data), "samples", tot)
    print("stalls:", ", ".join(f"{s[6:]} {100 * v / max(1, tot):.1f}%" for s, v in
                               sorted(agg.items(), key=lambda x: -x[1])[:9]))
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        t = r[idx["Source"]].split()
        o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        op[o] += int(r[S])
        ops[o] += int(r[idx["Instructions Executed"]])
    ti = sum(ops.values())
    print("opcode: samples% / executed%")
    for o, c in op.most_common(16):
        print(f"  {o:18s} {100 * c / max(1, tot):5.1f}%  {100 * ops[o] / max(1, ti):5.1f}%")
    print("hottest lines:")
    for r in sorted(data, key=lambda r: -int(r[S]))[:top]:
        st = sorted(((int(r[idx[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"  {int(r[S]):6d} {r[idx['Source']].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main()
