#!/usr/bin/env python
"""Front plan vs block plan on the same angles: max |diff| and norm error per case.
Usage: python tools/debug_front.py KIND:n:p[:S] ...   (developer tool, GPU box)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyramaterised_b200 as pyqc          # noqa: E402
from pyramaterised_b200 import engine      # noqa: E402

for w in sys.argv[1:]:
    f = w.split(":")
    kind, n, p = f[0], int(f[1]), int(f[2])
    S = int(f[3]) if len(f) > 3 else 1
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    ang = np.random.default_rng(n + p).random((S, max(1, qc.n_true_params))) * 2 * np.pi
    os.environ["PQC_FRONT"] = "1"
    st = qc.run_batch(ang)
    os.environ["PQC_FRONT"] = "0"
    st0 = qc.run_batch(ang)
    del os.environ["PQC_FRONT"]
    d = float((st - st0).abs().max().item())
    nrm = float(np.abs(engine.overlap(st, st).cpu().numpy() - 1).max())
    nrm0 = float(np.abs(engine.overlap(st0, st0).cpu().numpy() - 1).max())
    print(f"{w}: max|front-block| {d:.3e} norm_err front {nrm:.3e} block {nrm0:.3e}", flush=True)
    if d > 1e-10 and os.environ.get("DESCRIBE"):
        print("\n".join(l for l in qc.program.describe().split("\n") if "FRONT" in l), flush=True)
    del st, st0
    torch.cuda.empty_cache()
