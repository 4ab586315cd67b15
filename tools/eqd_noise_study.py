#!/usr/bin/env python
"""How often is the effective quantum dimension at cutoff 1e-12 (tests.py:211, BASELINE config 3)
decided by rounding noise?  TFIM 16q x 16 layers has a rank-16 QFIM of norm ~1e2, so its 17th
eigenvalue is pure noise of order 1e-12 -- in the reference as much as here.  This tool counts,
on the first K parameter sets of the bench workload, how the EQD moves between
  (a) the GPU path of bench.py (meet-in-the-middle plan, DMMA Gram, Jacobi eigenvalues),
  (b) the same QFIM with LAPACK eigenvalues on the host (isolates the eigen-solver),
  (c) the forward-only GPU plan (another summation order of the same mathematics),
  (d) the numpy oracle's QFIM (the reference's literal algorithm) + LAPACK, on M <= K sets,
and the EQD at cutoffs one decade either side.  Output: one JSON object (profiles/r2_eqd_noise.json).
Usage: python tools/eqd_noise_study.py [K] [M]"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np
import scipy.linalg
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                # noqa: E402
import pyramaterised_b200 as pyqc          # noqa: E402
from pyramaterised_b200 import engine      # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    n, p, cut = 16, 16, 1e-12
    qc = pyqc.templates.generate_circuit("TFIM", n, p)
    P = qc.n_true_params
    ang = np.random.default_rng(1).random((10000, P))[:K] * 2 * np.pi
    a = torch.from_numpy(ang).cuda()

    def eqds(w, c=cut):
        return (np.asarray(w) > c).sum(axis=1)

    Fa = qc.qfim_batch(a)
    wa = engine.eigvalsh(Fa).cpu().numpy()
    Fa = Fa.cpu().numpy()
    wb = np.stack([scipy.linalg.eigh(F, eigvals_only=True) for F in Fa])
    os.environ["PQC_BIDIR"] = "0"
    qc2 = pyqc.templates.generate_circuit("TFIM", n, p)        # planned without the meeting point
    Fc = qc2.qfim_batch(a)
    wc = engine.eigvalsh(Fc).cpu().numpy()
    Fc = Fc.cpu().numpy()
    del os.environ["PQC_BIDIR"]
    ea, eb, ec = eqds(wa), eqds(wb), eqds(wc)
    out = {"workload": f"TFIM {n}q x {p} layers, first {K} parameter sets of the bench stream",
           "cutoff": cut, "qfim_norm_median": float(np.median(np.abs(wa).max(axis=1))),
           "eqd_hist": {"gpu_bidir_jacobi": np.bincount(ea).tolist(),
                        "gpu_bidir_lapack": np.bincount(eb).tolist(),
                        "gpu_forward_jacobi": np.bincount(ec).tolist()},
           "eqd_differs_frac": {"jacobi_vs_lapack_same_qfim": float((ea != eb).mean()),
                                "bidir_vs_forward_plan": float((ea != ec).mean())},
           "qfim_max_rel_diff_bidir_vs_forward": float(np.abs(Fa - Fc).max() / np.abs(Fa).max()),
           "eigenvalue_17th_from_top": {
               "what": "sorted descending, index 16 (the first one past the rank)",
               "abs_quantiles_gpu": np.quantile(np.abs(np.sort(wa, axis=1)[:, ::-1][:, 16]),
                                                [0.05, 0.5, 0.95]).tolist(),
               "abs_quantiles_lapack": np.quantile(np.abs(np.sort(wb, axis=1)[:, ::-1][:, 16]),
                                                   [0.05, 0.5, 0.95]).tolist()},
           "eigenvalue_16th_from_top_min": float(np.sort(wa, axis=1)[:, ::-1][:, 15].min()),
           "eqd_hist_other_cutoffs": {str(c): np.bincount(eqds(wa, c)).tolist()
                                      for c in (1e-13, 1e-11, 1e-10, 1e-8)}}
    if M > 0:
        from oracle import pqc_oracle as orc
        specs, init = orc.generate_circuit("TFIM", n, p)
        cores = max(1, min(os.cpu_count() or 1, P + 1))
        eo, rel = [], []
        with mp.get_context("fork").Pool(cores) as pool:
            for i in range(M):
                e, F = bench.cpu_qfim_eqd(specs, n, ang[i], init, pool)
                eo.append(int(e))
                rel.append(float(np.abs(F - Fa[i]).max() / np.abs(F).max()))
        out["oracle"] = {"sets": M, "eqd_oracle": eo, "eqd_gpu": ea[:M].tolist(),
                         "eqd_differs_frac": float(np.mean(np.array(eo) != ea[:M])),
                         "qfim_max_rel_err": max(rel)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
