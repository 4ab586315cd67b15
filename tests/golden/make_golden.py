#!/usr/bin/env python
"""Generate tests/golden/ref_golden.npz by running the UNMODIFIED reference package.

Run in the authoring container only (it reads /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

The reference's only numerical backend (qutip==4.7.2, requirements.txt:15) cannot
be installed offline, so ``oracle/qutip_lite.py`` -- a restatement of the QuTiP
calls the reference makes -- is registered as ``qutip`` and the reference's own
gates.py / circuit.py / templates.py / measure.py then execute unchanged.  The
script first re-checks the reference's own known answers (tests.py:64-87,
114-128, 192-212, 284-295) so a broken shim cannot silently produce fixtures.
"""
import os
import sys
import time
from itertools import combinations

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import qutip_lite  # noqa: E402

qutip_lite.install_as_qutip()
sys.path.insert(0, "/root/reference")
import pyramaterised as ref  # noqa: E402
import qutip as qt  # noqa: E402  (the shim)

import cases  # noqa: E402


def vec(q):
    return np.asarray(q.full())[:, 0].copy()


def fresh_rng():
    """Put the reference's module-global generator (gates.py:10) back to seed 1."""
    ref.rng.bit_generator.state = np.random.default_rng(1).bit_generator.state


def self_check():
    c = ref.PQC(1)
    c.add_layer([ref.fixed_R_y(0, 1, np.pi / 2)])
    out = c.run("random")
    assert abs(np.real(out[1][0][0]) - 1 / np.sqrt(2)) < 1e-15
    c = ref.PQC(1)
    c.add_layer([ref.H(0, 1)], n=2)
    assert c.run("random") == qt.basis(2, 0)
    qg = cases.build_qg4(ref)
    e = qg.cost(cases.QG_ANGLES)
    assert abs(e - cases.QG_ENERGY) < 1e-5, e
    assert ref.measure.Measurements(qg).get_effective_quantum_dimension(1e-12) == cases.QG_EQD
    for N, P in ((4, 1), (4, 4), (6, 3), (6, 8)):
        layers, th = ref.templates.NPQC_layers(P, N)
        c = ref.PQC(N)
        for l in layers:
            c.add_layer(l)
        c.state = qt.Qobj(c.run(angles=th))
        Q = np.array(ref.measure.Measurements(c).get_QFI())
        assert np.abs(Q - np.eye(len(Q))).max() < 1e-12, (N, P)
    print("reference known answers reproduced on the shim")


def record_case(out, name):
    builder, S, G, want_magic = cases.CASES[name]
    t0 = time.time()
    qc = builder(ref)
    n = qc.n_qubits
    P = cases.n_true_params(qc)
    ang = cases.case_angles(name, qc, S)
    m = ref.measure.Measurements(qc)
    out[f"{name}/n"] = np.int64(n)
    out[f"{name}/P"] = np.int64(P)
    out[f"{name}/n_params_attr"] = np.int64(qc.n_params)
    out[f"{name}/parameterised"] = np.array(qc.parameterised, dtype=np.int64)
    out[f"{name}/angles"] = ang
    out[f"{name}/init"] = vec(qc.initial_state)
    states = [qc.run(list(a)) for a in ang]
    out[f"{name}/states"] = np.stack([vec(s) for s in states])
    out[f"{name}/cost"] = np.array([qc.cost(list(a)) for a in ang])
    out[f"{name}/Q"] = np.array([m.single_Q(s, n) for s in states])
    out[f"{name}/F"] = np.array([np.abs(a.overlap(b)) ** 2 for a, b in combinations(states, 2)])
    if want_magic:
        conv = m.get_conversion_matrices()
        out[f"{name}/renyi2"] = np.array([m.renyi_entropy_fast(s, conv) for s in states])
        out[f"{name}/gkp"] = np.array([m.gkp_fast(s, conv) for s in states])
    grads, qfis, eqds, nms, gvecs = [], [], [], [], []
    for a in ang[:G]:
        qc.update_state(list(a))
        gl = qc.get_gradients()
        grads.append(np.stack([vec(g) for g in gl]))
        F = m.get_QFI(grad_list=gl)
        qfis.append(F)
        eqds.append(m.get_effective_quantum_dimension(1e-12))
        nms.append(m.new_measure(F))
        gvecs.append(np.array(m.get_gradient_vector(list(a))))
    if G:
        out[f"{name}/grads"] = np.stack(grads)
        out[f"{name}/qfi"] = np.stack(qfis)
        out[f"{name}/eqd"] = np.array(eqds, dtype=np.int64)
        out[f"{name}/new_measure"] = np.array(nms)
        out[f"{name}/gradvec"] = np.stack(gvecs)
    print(f"  {name}: n={n} P={P} S={S} G={G}  {time.time() - t0:.1f}s")


def record_c1(out):
    """BASELINE config 1 exactly as the reference computes it: NPQC 4q/4 layers,
    expressibility(1000) then entanglement(1000) on the module RNG (quirk Q13)."""
    t0 = time.time()
    qc = ref.templates.generate_circuit("NPQC", 4, 4)
    m = ref.measure.Measurements(qc)
    fresh_rng()
    F = m._gen_f_samples(1000)                       # measure.py:123-137
    expr = m.expr(F, 2 ** 4)                         # measure.py:161-180
    prob, mid = m._gen_histo(F)
    counts, _ = np.histogram(F, bins=int((75 / 10000) * len(F)), range=(0, 1))
    ent = m.entanglement(1000)                       # measure.py:239-249 (next 1000 draws)
    F = np.array(F)
    out["c1/expr"] = np.float64(expr)
    out["c1/hist"] = counts.astype(np.int64)
    out["c1/prob_sum"] = np.float64(prob.sum())
    out["c1/F_head"] = F[:4096]
    out["c1/F_sum"] = np.float64(F.sum())
    out["c1/F_sqsum"] = np.float64((F ** 2).sum())
    out["c1/ent"] = np.array(ent)
    # a few alternative Hilbert-space sizes through expr(F, N) as find_eff_H uses it
    out["c1/expr_altN"] = np.array([m.expr(list(F), N) for N in (4, 8.5, 16, 64)])
    out["c1/expr_filt"] = np.float64(m.expr(list(F), 16, filt=0.2))
    print(f"  c1: expr={expr:.6f} mean Q={np.mean(ent):.6f}  {time.time() - t0:.1f}s")


def record_effm(out):
    """efficient_measurements (measure.py:370-459) on a 4-qubit HE circuit."""
    import random
    qc = ref.templates.generate_circuit("generic_HE", 4, 2)
    m = ref.measure.Measurements(qc)
    fresh_rng()
    d = m.efficient_measurements(40)
    out["effm/expr"] = np.float64(d["Expr"])
    out["effm/ent"] = np.array(d["Ent"])
    out["effm/magic"] = np.array(d["Magic"])
    out["effm/gkp"] = np.array(d["GKP"])
    fresh_rng()
    d = m.efficient_measurements(40, full_data=True)
    out["effm/full_expr"] = np.array(d["Expr"])
    out["effm/full_ent"] = np.array(d["Ent"])
    out["effm/full_magic"] = np.array(d["Magic"])
    out["effm/full_gkp"] = np.array(d["GKP"])
    random.seed(7)
    # fewer than 17 samples -> int(0.0075 * pairs) == 0 bins -> np.histogram raises
    # (measure.py:153-155); the replacement must raise the same ValueError.
    try:
        m.efficient_measurements(12, angles="clifford")
        raise AssertionError("reference should raise on zero bins")
    except ValueError:
        pass
    random.seed(7)
    d = m.efficient_measurements(20, angles="clifford")
    out["effm/cliff_magic"] = np.array(d["Magic"])
    out["effm/cliff_ent"] = np.array(d["Ent"])
    out["effm/cliff_expr"] = np.float64(d["Expr"])
    # 7 <= n < 12 branch: overlaps computed but KL skipped (measure.py:416-423)
    qc8 = ref.templates.generate_circuit("generic_HE", 7, 1)
    fresh_rng()
    d = ref.measure.Measurements(qc8).efficient_measurements(5, measure_eom=False,
                                                             measure_GKP=False)
    out["effm/n7_expr"] = np.float64(d["Expr"])
    out["effm/n7_ent"] = np.array(d["Ent"])
    out["effm/n7_magic"] = np.array(d["Magic"])
    print("  effm recorded")


def record_misc(out):
    # NPQC identity-QFIM at the reference angles, 8 qubits (tests.py:114-128)
    layers, th = ref.templates.NPQC_layers(3, 8)
    c = ref.PQC(8)
    for l in layers:
        c.add_layer(l)
    c.state = qt.Qobj(c.run(angles=th))
    out["npqc8/theta_ref"] = np.array(th, dtype=np.float64)
    out["npqc8/qfi"] = np.array(ref.measure.Measurements(c).get_QFI())
    # Bell state known answers (tests.py:284-295)
    bell = qt.states.bell_state("11")

    class _B:
        n_qubits = 2
        state = bell

        def run(self, a):
            return bell

    bm = ref.measure.Measurements(_B())
    out["bell/state"] = vec(bell)
    out["bell/vals"] = np.array([bm.renyi_entropy_fast(bell), bm.gkp_fast(bell),
                                 bm.single_Q(bell, 2)])
    # example.py path: expressibility(150) then entropy_of_magic(150) on the module RNG
    ex = cases.build_example4(ref)
    em = ref.measure.Measurements(ex)
    fresh_rng()
    out["example/expr150"] = np.float64(em.expressibility(150))
    out["example/eom150"] = np.float64(em.entropy_of_magic(150))
    print("  misc recorded")


def record_magic12(out):
    """BASELINE config 4 shape: one 12-qubit NPQC state through the reference's dense
    4096^3 ZGEMM formulation (measure.py:318-349)."""
    t0 = time.time()
    qc = ref.templates.generate_circuit("NPQC", 12, 3)
    P = cases.n_true_params(qc)
    ang = np.random.default_rng(4242).random((1, P)) * 2 * np.pi
    m = ref.measure.Measurements(qc)
    s = qc.run(list(ang[0]))
    conv = m.get_conversion_matrices()
    out["magic12/angles"] = ang
    out["magic12/state"] = vec(s)
    out["magic12/renyi2"] = np.float64(m.renyi_entropy_fast(s, conv))
    out["magic12/gkp"] = np.float64(m.gkp_fast(s, conv))
    out["magic12/Q"] = np.float64(m.single_Q(s, 12))
    print(f"  magic12: {out['magic12/renyi2']:.9f} {out['magic12/gkp']:.9f}  "
          f"{time.time() - t0:.1f}s")


def main():
    self_check()
    out = {}
    for name in sorted(cases.CASES):
        record_case(out, name)
    record_c1(out)
    record_effm(out)
    record_misc(out)
    if "--no-12q" not in sys.argv:
        record_magic12(out)
    path = os.path.join(HERE, "ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB, {len(out)} arrays")


if __name__ == "__main__":
    main()
