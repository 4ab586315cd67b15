#!/usr/bin/env python
"""Fixtures for the last two fenced code paths -> tests/golden/ref_golden_r3.npz.

Same recipe as make_golden.py / make_golden_r2.py (the UNMODIFIED reference package from
/root/reference running on oracle/qutip_lite.py registered as ``qutip``; authoring container only):

  * gates.py:458-466   shared_parameter(commute=False) with members that do not commute: states,
                       derivative states (circuit.py:149-192), QFIM (measure.py:33-71), EQD, cost
  * gates.py:75-85     Gate.__add__ / __radd__: dense sums and their action on a state

    python tests/golden/make_golden_r3.py
"""
import os
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

import cases_r3  # noqa: E402


def main():
    out = {}
    for k, ang in enumerate(cases_r3.NONCOMM3_ANGLES):
        qc = cases_r3.build_noncommuting3(ref)
        m = ref.measure.Measurements(qc)
        st = qc.update_state(list(ang))
        out[f"noncomm3/{k}/state"] = np.asarray(st.full())[:, 0]
        out[f"noncomm3/{k}/cost"] = np.float64(qc.cost(list(ang)))
        grads = qc.get_gradients()
        out[f"noncomm3/{k}/grads"] = np.stack([np.asarray(g.full())[:, 0] for g in grads])
        out[f"noncomm3/{k}/qfi"] = np.asarray(m.get_QFI(), dtype=np.float64)
        out[f"noncomm3/{k}/eqd"] = np.int64(m.get_effective_quantum_dimension(1e-12))
        print(f"  noncomm3 {k}: cost {out[f'noncomm3/{k}/cost']:.9f} eqd {out[f'noncomm3/{k}/eqd']}"
              f" |grads| {np.abs(out[f'noncomm3/{k}/grads']).max():.4f}")
    N, a, b, c = cases_r3.build_sum_gates(ref)
    psi = ref.PQC(N)
    psi.add_layer([ref.R_y(i, N) for i in range(N)] + [ref.CHAIN(ref.CNOT, N)])
    st = psi.run([0.4, 1.3, 2.2])
    out["sum/state_in"] = np.asarray(st.full())[:, 0]
    out["sum/a_plus_b"] = np.asarray((a + b).full())
    out["sum/a_plus_b_plus_c"] = np.asarray(((a + b) + c).full())
    out["sum/radd"] = np.asarray((a.operation + c).full())       # Qobj + Gate -> Gate.__radd__
    out["sum/applied"] = np.asarray(((a + b) * st).full())[:, 0]
    path = os.path.join(HERE, "ref_golden_r3.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e3:.1f} kB, {len(out)} arrays")


if __name__ == "__main__":
    main()
