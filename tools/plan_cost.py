#!/usr/bin/env python
"""Offline cost proxy of the front planner's plans (planning is host-only): parses
Program.describe() and adds up, in units of one 4-slot layer op on a 4096-amplitude tile,
  pass: max(HBM round trip ~ 8 units at 16q, sum over sweeps of [0.9 smem round trip + ops]).
Usage: python tools/plan_cost.py KIND:n:p ...   (env PQC_FRONT_ALPHA etc. apply)"""
import os
import re
import sys
from collections import Counter

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyramaterised_b200 as pyqc          # noqa: E402

COST = {"35": 1.0, "36": 1.0, "37": 1.0, "38": 1.0, "32": 0.6, "33": 0.3, "34": 0.6, "8": 0.12, "2": 0.12,
        "39": 0.02, "40": 0.1, "4": 0.05, "7": 0.05}


def plan_cost(kind, n, p):
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    tot, nops, nsw, npass = 0.0, Counter(), 0, 0
    for l in qc.program.describe().split("\n"):
        if not l.startswith("FRONT PASS"):
            continue
        npass += 1
        comp = 0.0
        for rb, ops in re.findall(r"\[rb ([0-9,]+) pre\d+ post\d+:([0-9 ]*)\]", l):
            nsw += 1
            comp += 0.9
            for o in ops.split():
                comp += COST.get(o, 0.5)
                nops[o] += 1
        tot += max(8.0, comp)
    return tot, npass, nsw, dict(nops)


if __name__ == "__main__":
    for w in sys.argv[1:]:
        k, n, p = w.split(":")
        c, npass, nsw, ops = plan_cost(k, int(n), int(p))
        l4 = sum(v for o, v in ops.items() if o in ("35", "36", "37", "38"))
        print(f"{w}: cost {c:.1f} passes {npass} sweeps {nsw} layer-ops {l4} ops {ops}")
