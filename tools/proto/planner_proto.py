#!/usr/bin/env python
"""Prototype of the round-2 'front' planner (host-only, no GPU): DAG readiness + greedy
tile / sweep selection.  Prints passes / sweeps / ops per sweep for the target circuits so the
strategy can be judged before it is written in C++ (csrc/pqc_front.cu)."""
import itertools
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pyramaterised_b200 as pyqc
from pyramaterised_b200 import _lib as L

RX, RY, RZ, H, X, S, T, CNOT, CZ, SQ, RXX, RYY, RZZ, FSIM, FFSIM, IDENT = range(16)
ZZSUM, RXY = 32, 33


class Op:
    __slots__ = ("kind", "b0", "b1", "param", "group", "support", "mix", "diag", "pauli", "px", "pz",
                 "two", "px2", "pz2", "pairs", "tgt", "ctl", "scale", "offset")


def make(op, n):
    kind, q0, q1, param, param2, group, scale, offset = op
    o = Op()
    o.kind, o.param, o.group, o.scale, o.offset = kind, param, group, scale, offset
    o.b0 = n - 1 - q0
    o.b1 = n - 1 - q1 if q1 >= 0 else -1
    m0 = 1 << o.b0
    m1 = (1 << o.b1) if o.b1 >= 0 else 0
    o.support = m0 | m1
    o.mix = 0; o.diag = False; o.pauli = False; o.px = o.pz = 0; o.two = False; o.px2 = o.pz2 = 0
    o.pairs = []; o.tgt = 0; o.ctl = 0
    if kind == RX: o.mix = m0; o.pauli = True; o.px = m0
    elif kind == RY: o.mix = m0; o.pauli = True; o.px = m0; o.pz = m0
    elif kind == RZ: o.diag = True; o.pauli = True; o.pz = m0
    elif kind == H: o.mix = m0
    elif kind == X: o.tgt = m0
    elif kind in (S, T, IDENT): o.diag = True
    elif kind == CNOT: o.tgt = m1; o.ctl = m0
    elif kind == CZ: o.diag = True
    elif kind == RXX: o.mix = m0 | m1; o.pauli = True; o.px = m0 | m1
    elif kind == RYY: o.mix = m0 | m1; o.pauli = True; o.px = m0 | m1; o.pz = m0 | m1
    elif kind == RZZ: o.diag = True; o.pauli = True; o.pz = m0 | m1
    else: o.mix = m0 | m1
    return o


def pc(x1, z1, x2, z2):
    return ((bin(x1 & z2).count("1") + bin(z1 & x2).count("1")) & 1) == 0


def commute(a, b):
    if not (a.support & b.support): return True
    if a.diag and b.diag: return True
    if a.pauli and b.pauli:
        ok = pc(a.px, a.pz, b.px, b.pz)
        if a.two: ok = ok and pc(a.px2, a.pz2, b.px, b.pz)
        if b.two: ok = ok and pc(a.px, a.pz, b.px2, b.pz2)
        if a.two and b.two: ok = ok and pc(a.px2, a.pz2, b.px2, b.pz2)
        return ok
    # permutation ops (X, CNOT): diagonal in the control's Z basis, X-type on the target
    for p, o in ((a, b), (b, a)):
        if p.kind in (CNOT, X):
            if o.kind in (CNOT, X):
                return not (p.tgt & o.ctl) and not (p.ctl & o.tgt)
            if o.diag: return not (o.support & p.tgt)
            if o.pauli and not o.two: return not (o.support & p.ctl) and not (o.pz & p.tgt)
            return False
    return False


def fuse(ops):
    out = []
    i = 0
    while i < len(ops):
        j = i
        while j < len(ops) and ops[j].group == ops[i].group: j += 1
        run = ops[i:j]
        i = j
        if len(run) >= 2 and all(o.pauli for o in run) and all(commute(a, b) for a, b in itertools.combinations(run, 2)):
            dead = [False] * len(run)
            for a in range(len(run)):
                if True or dead[a] or run[a].kind != RZZ: continue   # ZZ stays per bond (merged per sweep at emission)
                mem = [b for b in range(a, len(run)) if not dead[b] and run[b].kind == RZZ and run[b].param == run[a].param]
                if len(mem) < 2 or len({run[b].support for b in mem}) < len(mem): continue
                z = run[a]; z.kind = ZZSUM; z.pauli = False; z.diag = True; z.mix = 0; z.support = 0
                for b in mem:
                    z.pairs.append((run[b].b0, run[b].b1)); z.support |= (1 << run[b].b0) | (1 << run[b].b1)
                    if b != a: dead[b] = True
            for a in range(len(run)):
                if dead[a] or run[a].kind != RYY: continue
                for b in range(len(run)):
                    if dead[b] or run[b].kind != RXX or run[b].param != run[a].param or run[b].support != run[a].support: continue
                    run[a].kind = RXY; run[a].two = True; run[a].px2 = run[b].px; run[a].pz2 = run[b].pz; dead[b] = True
                    break
            run = [o for o, d in zip(run, dead) if not d]
        out += run
    return out


def build_dag(ops):
    N = len(ops)
    preds = [[] for _ in range(N)]
    # per-bit list of earlier ops touching the bit keeps the pair scan short
    for j in range(N):
        for i in range(j - 1, -1, -1):
            if (ops[i].support & ops[j].support) and not commute(ops[i], ops[j]):
                preds[j].append(i)
    return preds


def simulate(ops, preds, done, tile, R, limit=None):
    """ops executable in one sweep with register bit set R (mask) inside tile (mask): returns the
    list of op indices in execution order.  R = None: unlimited sweeps (tile-level light cone)."""
    newly = []
    mark = set()
    changed = True
    # candidate frontier: ops not done; iterate in index order until fixpoint
    pending = [i for i in range(len(ops)) if not done[i]]
    while changed:
        changed = False
        rest = []
        for i in pending:
            o = ops[i]
            if any(not (done[p] or p in mark) for p in preds[i]):
                rest.append(i); continue
            ok = True
            if o.mix:
                ok = (o.mix & ~(tile if R is None else R)) == 0
            elif o.tgt:
                ok = (o.tgt & ~tile) == 0
            if ok:
                mark.add(i); newly.append(i); changed = True
            else:
                rest.append(i)
        pending = rest
    return newly


def weight(ops, idxs):
    return sum(1 for i in idxs if ops[i].mix)


def plan(ops, n, verbose=False, max_sweeps=24, min_w=1):
    preds = build_dag(ops)
    done = [False] * len(ops)
    low = (1 << min(4, n)) - 1
    passes = []
    while not all(done):
        # ---- tile: grow greedily by light-cone gain per added bit; candidates are the missing bits
        # of pending mixing ops (1 or 2 bits at a time, so 2-qubit gates can seed a window)
        tile = low
        cap = min(12, n)
        while bin(tile).count("1") < cap:
            base = weight(ops, simulate(ops, preds, done, tile, None))
            cands = set()
            for i, o in enumerate(ops):
                if done[i] or not (o.mix or o.tgt): continue
                miss = (o.mix | o.tgt) & ~tile
                if miss and bin(miss).count("1") + bin(tile).count("1") <= cap: cands.add(miss)
            if not cands: break
            best, bm = -1.0, None
            for miss in sorted(cands):
                g = weight(ops, simulate(ops, preds, done, tile | miss, None)) - base
                score = g / bin(miss).count("1")
                if score > best: best, bm = score, miss
            tile |= bm
        for b in range(n):                      # fill up
            if bin(tile).count("1") >= cap: break
            tile |= 1 << b
        tbits = [b for b in range(n) if (tile >> b) & 1]
        # ---- sweeps
        sweeps = []
        while len(sweeps) < max_sweeps:
            best, bestR, bestl = -1, None, None
            for comb in itertools.combinations(tbits, 4):
                R = sum(1 << b for b in comb)
                l = simulate(ops, preds, done, tile, R)
                w = weight(ops, l)
                if w > best: best, bestR, bestl = w, R, l
            if best < min_w and sweeps: break
            if best <= 0 and not bestl: break
            for i in bestl: done[i] = True
            sweeps.append((bestR, bestl))
            if best <= 0: break
        if not sweeps:
            raise RuntimeError("no progress")
        passes.append((tile, sweeps))
        if verbose:
            print(f"pass {len(passes)}: tile {tile:0{n}b} sweeps {len(sweeps)}: " +
                  " ".join(f"[{weight(ops, l)}m/{len(l)}]" for R, l in sweeps))
    return passes


if __name__ == "__main__":
    kind, n, p = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False) if kind == "XXZ" else pyqc.templates.generate_circuit(kind, n, p)
    ops = fuse([make(o, n) for o in qc.lower()])
    print(kind, n, p, "ops", len(ops), "mixing", sum(1 for o in ops if o.mix))
    ps = plan(ops, n, verbose=True, min_w=int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    print("passes", len(ps), "sweeps", sum(len(s) for _, s in ps))
