#!/usr/bin/env python
"""Round-2 additions to the reference-generated fixtures -> tests/golden/ref_golden_r2.npz.

Same recipe as make_golden.py (the UNMODIFIED reference package from /root/reference running
on oracle/qutip_lite.py registered as ``qutip``; authoring container only):

  * measure.py:268-316  get_conversion_matrices (xor / sign tables), n = 1, 3, 5
  * measure.py:101-121  find_overparam_point on a 3-qubit circuit, module RNG at seed 1
  * measure.py:473-553  train(method="QNG") and train(method="gradient") on a 3-qubit TFIM,
                        fixed start angles: energies, trajectories, magic / Q / GKP traces
  * measure.py:199-224  find_eff_H on fidelity samples of a half-filled zfsim circuit and the
                        fidelity samples themselves (the zfsim half of tests.py:311-342)
  * gates.py:407-435    ARBGATE circuits (dense expm on the shim): states, derivative states,
                        QFIM, EQD, cost
  * full-depth rows used by the GPU tests at sizes the oracle cannot reach with QuTiP's dense
    operators are NOT generated here (the numpy oracle covers those).

    python tests/golden/make_golden_r2.py
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import qutip_lite  # noqa: E402

qutip_lite.install_as_qutip()
sys.path.insert(0, "/root/reference")
import pyramaterised as ref  # noqa: E402

import cases_r2  # noqa: E402


def fresh_rng():
    ref.rng.bit_generator.state = np.random.default_rng(1).bit_generator.state


def main():
    out = {}
    # ---- conversion matrices
    for n in (1, 3, 5):
        c = ref.PQC(n)
        m = ref.measure.Measurements(c)
        xor, sign = m.get_conversion_matrices()
        out[f"conv/{n}/xor"] = np.asarray(xor, dtype=np.int64)
        out[f"conv/{n}/sign"] = np.asarray(sign, dtype=np.int64)
        out[f"conv/{n}/base3"] = np.asarray(m.numberToBase(3 % (2 ** n), 2, n), dtype=np.int64)
    # ---- find_overparam_point
    qc = cases_r2.build_overparam3(ref)
    m = ref.measure.Measurements(qc)
    fresh_rng()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        count = m.find_overparam_point([0])
    out["overparam/count"] = np.int64(count)
    out["overparam/log"] = np.array(buf.getvalue())
    out["overparam/n_layers_after"] = np.int64(qc.n_layers)
    # ---- train: QNG and plain gradient descent
    for method, rate, eps in (("QNG", 0.05, 1e-4), ("gradient", 0.05, 1e-5)):
        qc = cases_r2.build_tfim3(ref)
        m = ref.measure.Measurements(qc)
        energy, traj, magics, ents, gkps = m.train(epsilon=eps, rate=rate, method=method,
                                                   angles=list(cases_r2.TFIM3_START))
        out[f"train/{method}/energy"] = np.float64(energy)
        out[f"train/{method}/traj"] = np.array(traj, dtype=np.float64)
        out[f"train/{method}/magics"] = np.array(magics, dtype=np.float64)
        out[f"train/{method}/ents"] = np.array(ents, dtype=np.float64)
        out[f"train/{method}/gkps"] = np.array(gkps, dtype=np.float64)
        out[f"train/{method}/final_angles"] = np.array(qc.get_params(), dtype=np.float64)
        print(f"  train {method}: {len(traj)} points, E = {energy:.9f}")
    # ---- zfsim half of test_effective_hilbert_space: F samples + the fitted dimension
    random.seed(11)
    for n in (4, 6):
        qc = ref.templates.generate_circuit("zfsim", n, n)
        m = ref.measure.Measurements(qc)
        fresh_rng()
        F = m._gen_f_samples(60)
        out[f"zfsim/{n}/init"] = np.asarray(qc.initial_state.full())[:, 0]
        out[f"zfsim/{n}/F"] = np.array(F, dtype=np.float64)
        out[f"zfsim/{n}/effH"] = np.float64(m.find_eff_H(F, n))
        print(f"  zfsim {n}: eff_H = {out[f'zfsim/{n}/effH']:.4f}")
    # ---- ARBGATE: states, derivative states, QFIM, cost through the reference's own classes
    for k, ang in enumerate(cases_r2.ARB4_ANGLES):
        qc = cases_r2.build_arb4(ref)
        m = ref.measure.Measurements(qc)
        st = qc.run(list(ang))
        out[f"arb4/{k}/state"] = np.asarray(st.full())[:, 0]
        out[f"arb4/{k}/cost"] = np.float64(qc.cost(list(ang)))
        grads = qc.get_gradients()
        out[f"arb4/{k}/grads"] = np.stack([np.asarray(g.full())[:, 0] for g in grads])
        out[f"arb4/{k}/qfi"] = np.asarray(m.get_QFI(), dtype=np.float64)
        out[f"arb4/{k}/eqd"] = np.int64(m.get_effective_quantum_dimension(1e-12))
        print(f"  arb4 {k}: cost {out[f'arb4/{k}/cost']:.9f} eqd {out[f'arb4/{k}/eqd']}")
    path = os.path.join(HERE, "ref_golden_r2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e3:.1f} kB, {len(out)} arrays")


if __name__ == "__main__":
    main()
