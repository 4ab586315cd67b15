#!/usr/bin/env python
"""Quick A/B timing of the hot kernels (developer tool, not the bench.py contract).

  PQC_LIB_PATH=pyramaterised_b200/variants/libX.so python tools/microbench.py [--circuit TFIM]

Prints one JSON line: pure gate application (run_batch) and QFIM throughput on a small batch.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyramaterised_b200 as pyqc          # noqa: E402
from pyramaterised_b200 import engine      # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--circuit", default="TFIM")
    ap.add_argument("--qubits", type=int, default=16)
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--states", type=int, default=4096)
    ap.add_argument("--sets", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tag", default=os.environ.get("PQC_LIB_PATH", "default"))
    a = ap.parse_args()
    kw = {"shuffle": False} if a.circuit == "XXZ" else {}
    qc = pyqc.templates.generate_circuit(a.circuit, a.qubits, a.layers, **kw)
    P = qc.n_true_params
    rng = np.random.default_rng(1)
    ang_run = torch.from_numpy(rng.random((a.states, P)) * 2 * np.pi).cuda()
    ang_q = torch.from_numpy(rng.random((a.sets, P)) * 2 * np.pi).cuda()
    out = {"tag": a.tag, "circuit": f"{a.circuit}-{a.qubits}x{a.layers}"}
    buf = torch.empty((a.states, 1 << a.qubits), dtype=torch.complex128, device="cuda")
    engine.profile_begin()
    ms = timed(lambda: qc.program.run(ang_run, out=buf), a.reps)
    prof = engine.profile_end()
    out["run_states_per_s"] = a.states / (ms / 1e3)
    out["run_pass_GBps"] = prof["bytes"] / (prof["ms"] / 1e3) / 1e9 if prof["ms"] > 0 else 0
    engine.profile_begin()
    ms = timed(lambda: qc.program.qfim(ang_q), a.reps)
    prof = engine.profile_end()
    out["qfim_sets_per_s"] = a.sets / (ms / 1e3)
    out["qfim_pass_GBps"] = prof["bytes"] / (prof["ms"] / 1e3) / 1e9 if prof["ms"] > 0 else 0
    out["qfim_pass_share"] = prof["ms"] / (ms * (a.reps + 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
