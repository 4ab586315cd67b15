#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs on one B200 (not the bench.py contract).
Writes one JSON object per config to stdout; used to fill profiles/r1_configs.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyramaterised_b200 as pyqc          # noqa: E402
from pyramaterised_b200 import engine      # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def reseed():
    pyqc.gates.rng.bit_generator.state = np.random.default_rng(1).bit_generator.state


def c1():
    qc = pyqc.templates.generate_circuit("NPQC", 4, 4)
    m = pyqc.measure.Measurements(qc)

    def run():
        reseed()
        return m.expressibility(1000), np.mean(m.entanglement(1000))
    ms, (e, q) = timed(run)
    return {"config": "C1 NPQC 4q x 4 layers, expressibility + entanglement, S=1000", "ms": ms,
            "samples_per_s": 2000 / (ms / 1e3), "expr": e, "mean_Q": float(q)}


def c2(S):
    qc = pyqc.templates.generate_circuit("generic_HE", 10, 10)
    P = qc.n_true_params
    ang = torch.from_numpy(np.random.default_rng(1).random((S, P)) * 2 * np.pi).cuda()
    ms_run, st = timed(lambda: qc.run_batch(ang))
    ms_q, Q = timed(lambda: engine.meyer_wallach(st))
    pairs = S * (S - 1) // 2
    bins = engine.n_bins(pairs)
    ms_f, hist = timed(lambda: engine.fidelity_hist(st, bins=bins)[0], reps=1)
    ms_k, kl = timed(lambda: engine.kl_haar(hist, 2.0 ** 10))
    return {"config": f"C2 generic_HE 10q x 10 layers, expressibility + entanglement, S={S}",
            "ms_states": ms_run, "states_per_s": S / (ms_run / 1e3),
            "apply_GBps_algorithmic": 11 * 2 * 16 * 1024 * S / (ms_run / 1e3) / 1e9,
            "ms_meyer_wallach": ms_q, "ms_pair_hist": ms_f, "pairs": pairs, "bins": bins,
            "pair_Tflops": 8 * 1024 * pairs / (ms_f / 1e3) / 1e12, "ms_kl": ms_k,
            "expr": float(kl.item()), "mean_Q": float(Q.mean().item()),
            "hist_total": int(hist.sum().item())}


def c3_apply(kind="TFIM", n=16, p=16, S=4096):
    """Pure state generation (PQC.run over a batch) at the headline size: the gate-apply
    roofline of SURVEY 8d, L * 2 * 16 * 2^n bytes per state with L template layers."""
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    ang = torch.from_numpy(np.random.default_rng(1).random((S, qc.n_true_params)) * 2 * np.pi).cuda()
    out = torch.empty((S, 2 ** n), dtype=torch.complex128, device="cuda")
    engine.profile_begin()
    ms, _ = timed(lambda: qc.program.run(ang, init=qc.initial_state.tensor, out=out))
    prof = engine.profile_end()
    L = p + 1 if kind in ("TFIM", "NPQC", "generic_HE") else p
    return {"config": f"C3 apply-only {kind} {n}q x {p} layers, S={S}", "ms": ms,
            "states_per_s": S / (ms / 1e3), "passes": qc.program.n_passes,
            "algorithmic_GBps_layers": L * 2 * 16 * 2 ** n * S / (ms / 1e3) / 1e9,
            "pass_kernel_GBps": prof["bytes"] / (prof["ms"] / 1e3) / 1e9,
            "pass_kernel_launches": prof["launches"]}


def c4(S=1000, p=12):
    qc = pyqc.templates.generate_circuit("NPQC", 12, p)
    ang = torch.from_numpy(np.random.default_rng(1).random((S, qc.n_true_params)) * 2 * np.pi).cuda()
    ms_run, st = timed(lambda: qc.run_batch(ang))
    ms_m, mg = timed(lambda: engine.magic(st, (2.0, 0.5)), reps=2)
    return {"config": f"C4 NPQC 12q x {p} layers, Renyi-2 magic + GKP, S={S}", "ms_states": ms_run,
            "ms_magic": ms_m, "samples_per_s": S / ((ms_run + ms_m) / 1e3),
            "magic_Gflops_fwht": S * 4096 * 4096 * 12 / (ms_m / 1e3) / 1e9,
            "mean_magic": float(mg[0].mean().item()),
            "mean_gkp": float((mg[1] / (2 * np.log(2))).mean().item())}


def c5(n=28, p=20, S=2):
    qc = pyqc.templates.generate_circuit("NPQC", n, p)
    ang = torch.from_numpy(np.random.default_rng(1).random((S, qc.n_true_params)) * 2 * np.pi).cuda()
    ms_run, st = timed(lambda: qc.run_batch(ang), reps=1)
    nrm = engine.overlap(st, st).cpu().numpy()
    ms_q, Q = timed(lambda: engine.meyer_wallach(st), reps=1)
    return {"config": f"C5 probe NPQC {n}q x {p} layers, S={S} ({16 * 2 ** n / 2 ** 30:.0f} GiB per state)",
            "ms_states": ms_run, "passes": qc.program.n_passes,
            "apply_GBps_algorithmic": p * 2 * 16 * 2 ** n * S / (ms_run / 1e3) / 1e9,
            "norm_err": float(np.abs(nrm - 1).max()), "ms_meyer_wallach": ms_q,
            "Q": Q.cpu().numpy().tolist()}


def c5x(n=28, p=20, S=17, block=6):
    """Config 5 shape end to end at a sample count that finishes in seconds: block-streamed
    expressibility + entanglement (two blocks of `block` 4 GiB states resident)."""
    qc = pyqc.templates.generate_circuit("NPQC", n, p)
    m = pyqc.measure.Measurements(qc)
    reseed()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e, Q = m.expressibility_streamed(S, block, want_Q=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nb = (S + block - 1) // block
    gens = sum(min(S, (i + 1) * block) - i * block + sum(min(S, (j + 1) * block) - j * block
               for j in range(i + 1, nb)) for i in range(nb))
    return {"config": f"C5 streamed NPQC {n}q x {p} layers, S={S}, block={block}: expressibility + entanglement",
            "seconds": dt, "state_generations": gens, "pairs": S * (S - 1) // 2,
            "bins": engine.n_bins(S * (S - 1) // 2), "expr": e, "mean_Q": float(np.mean(Q))}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]
    out = []
    for w in which:
        if w == "c1": r = c1()
        elif w.startswith("c2"): r = c2(int(w.split(":")[1]) if ":" in w else 100000)
        elif w == "c4": r = c4()
        elif w.startswith("c3"):                      # c3[:KIND[:n[:p[:S]]]]
            f = w.split(":")
            r = c3_apply(f[1] if len(f) > 1 else "TFIM", int(f[2]) if len(f) > 2 else 16,
                         int(f[3]) if len(f) > 3 else 16, int(f[4]) if len(f) > 4 else 4096)
        elif w.startswith("c5x"): r = c5x(int(w.split(":")[1]) if ":" in w else 28)
        elif w.startswith("c5"): r = c5(int(w.split(":")[1]) if ":" in w else 28)
        print(json.dumps(r), flush=True)
