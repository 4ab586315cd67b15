#!/usr/bin/env python
"""Developer tool: QFIM of one circuit vs the numpy oracle under every engine knob (each in its
own process: most knobs are read once).  python tools/debug_qfim.py KIND N P [knob=val ...]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KNOBS = [{}, {"PQC_BIDIR": "0"}, {"PQC_BIDIR_TRAIL": "0"}, {"PQC_QFIM_GRAM": "0"},
         {"PQC_FAST": "0"}, {"PQC_ENGINE": "v0"}, {"PQC_BIDIR": "0", "PQC_FAST": "0"}]


def child(kind, n, p):
    import pyramaterised_b200 as pyqc
    from oracle import pqc_oracle as orc
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    specs, init = orc.generate_circuit(kind, n, p)
    ang = np.random.default_rng(n + p).random((2, orc.n_params(specs))) * 2 * np.pi
    F, st = qc.qfim_batch(ang, want_states=True)
    F = F.cpu().numpy()
    ref_st = orc.run(specs, n, ang, init)
    gr = orc.gradients(specs, n, ang, init)
    print("  state err %.2e" % np.abs(st.cpu().numpy() - ref_st).max())
    for s in range(2):
        R = orc.qfi(ref_st[s], gr[s])
        E = np.abs(F[s] - R)
        print("  set %d: max abs err %.3e (|F| %.3g)" % (s, E.max(), np.abs(R).max()))
        if E.max() > 1e-8 and R.shape[0] <= 8:
            np.set_printoptions(precision=4, linewidth=200, suppress=True)
            print("  got\n", F[s], "\n  ref\n", R)
        elif E.max() > 1e-8:
            bad = np.argwhere(E > 1e-8)
            print("  bad entries:", len(bad), "rows", sorted(set(bad[:, 0].tolist())))
    g = qc.program.gradients(ang, init=qc.initial_state.tensor).cpu().numpy()
    print("  gradient-state err %.2e" % np.abs(g[:, 1:] - gr).max(), "per parameter:",
          " ".join("%.1e" % np.abs(g[:, 1 + q] - gr[:, q]).max() for q in range(gr.shape[1])))


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    kind, n, p = sys.argv[1], sys.argv[2], sys.argv[3]
    for kn in KNOBS:
        env = dict(os.environ)
        env.update(kn)
        print(f"== {kind} {n} {p} {kn}", flush=True)
        r = subprocess.run([sys.executable, __file__, "--child", kind, n, p], env=env,
                           capture_output=True, text=True)
        print(r.stdout, r.stderr[-800:] if r.returncode else "", flush=True)
