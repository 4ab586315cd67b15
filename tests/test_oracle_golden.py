"""Pin oracle/pqc_oracle.py against fixtures produced by the unmodified reference
(tests/golden/make_golden.py) and the reference's own known answers (tests.py)."""
import numpy as np
import pytest

import cases
from helpers import oracle_case
from oracle import pqc_oracle as orc

ALL = sorted(cases.CASES)
ATOL = 1e-10          # amplitudes / fidelities (BASELINE.json north_star)
RTOL = 1e-8           # QFIM / magic


@pytest.mark.parametrize("name", ALL)
def test_states_cost_Q_F(golden, name):
    specs, n, init = oracle_case(name)
    ang = golden[f"{name}/angles"]
    assert orc.n_params(specs) == int(golden[f"{name}/P"]) == ang.shape[1]
    par, tot = orc.parameterised_attr(specs)
    assert tot == int(golden[f"{name}/n_params_attr"])               # quirk Q1
    assert par == list(golden[f"{name}/parameterised"])
    if init is not None:
        assert np.array_equal(init, golden[f"{name}/init"])
    st = orc.run(specs, n, ang, init)
    assert np.abs(st - golden[f"{name}/states"]).max() < ATOL
    assert np.abs(orc.cost_zz(st) - golden[f"{name}/cost"]).max() < ATOL
    assert np.abs([orc.single_Q(s, n) for s in st] - golden[f"{name}/Q"]).max() < ATOL
    assert np.abs(orc.fidelity_samples(st) - golden[f"{name}/F"]).max() < ATOL


@pytest.mark.parametrize("name", [c for c in ALL if cases.CASES[c][3]])
def test_magic(golden, name):
    st = golden[f"{name}/states"]
    for s, r2, g in zip(st, golden[f"{name}/renyi2"], golden[f"{name}/gkp"]):
        assert abs(orc.renyi_dense(s, 2.0) - r2) < RTOL * max(1, abs(r2))
        assert abs(orc.renyi_fwht(s, 2.0) - r2) < RTOL * max(1, abs(r2))
        assert abs(orc.gkp(s) - g) < RTOL * max(1, abs(g))
        assert abs(orc.renyi_fwht(s, 0.5) / (2 * np.log(2)) - g) < RTOL * max(1, abs(g))


@pytest.mark.parametrize("name", [c for c in ALL if cases.CASES[c][2] > 0])
def test_gradients_qfi_eqd(golden, name):
    specs, n, init = oracle_case(name)
    G = cases.CASES[name][2]
    ang = golden[f"{name}/angles"][:G]
    st = orc.run(specs, n, ang, init)
    gr = orc.gradients(specs, n, ang, init)
    assert np.abs(gr - golden[f"{name}/grads"]).max() < ATOL
    for s in range(G):
        F = orc.qfi(st[s], gr[s])
        ref = golden[f"{name}/qfi"][s]
        assert np.abs(F - ref).max() < RTOL * max(1.0, np.abs(ref).max())
        assert orc.eqd(F, 1e-12) == int(golden[f"{name}/eqd"][s])
        assert abs(orc.new_measure(F) - golden[f"{name}/new_measure"][s]) < 1e-8


def test_reference_known_answers(golden):
    # tests.py:192-212
    specs, n, _ = oracle_case("qg4")
    st = orc.run(specs, n, [cases.QG_ANGLES])
    assert abs(orc.cost_zz(st)[0] - cases.QG_ENERGY) < 1e-5
    gr = orc.gradients(specs, n, [cases.QG_ANGLES])
    assert orc.eqd(orc.qfi(st[0], gr[0]), 1e-12) == cases.QG_EQD
    # tests.py:64-87
    v = orc.run([("fixed_R_y", 0, np.pi / 2)], 1, np.zeros((1, 0)))
    assert abs(v[0, 1].real - 1 / np.sqrt(2)) < 1e-15
    v = orc.run([("H", 0), ("H", 0)], 1, np.zeros((1, 0)))
    assert np.abs(v[0] - [1, 0]).max() < 1e-12
    # tests.py:284-295
    bell = golden["bell/state"]
    assert abs(orc.renyi(bell)) < 1e-10 and abs(orc.gkp(bell)) < 1e-10
    assert abs(orc.single_Q(bell, 2) - 1) < 1e-10
    assert np.abs(golden["bell/vals"] - [0, 0, 1]).max() < 1e-10


@pytest.mark.parametrize("N,P", [(4, 1), (4, 4), (6, 3), (6, 8), (8, 3)])
def test_npqc_identity_qfim(golden, N, P):
    # tests.py:114-128
    specs, th = orc.npqc(P, N)
    st = orc.run(specs, N, [th])
    F = orc.qfi(st[0], orc.gradients(specs, N, [th])[0])
    assert np.abs(F - np.eye(len(F))).max() < 1e-12
    if (N, P) == (8, 3):
        assert np.allclose(th, golden["npqc8/theta_ref"])
        assert np.abs(F - golden["npqc8/qfi"]).max() < 1e-12


def test_c1_expressibility_entanglement(golden):
    """BASELINE config 1 with the reference's own RNG stream (gates.py:10, quirk Q13)."""
    specs, _ = orc.generate_circuit("NPQC", 4, 4)
    P = orc.n_params(specs)
    draws = np.random.default_rng(1).random(2 * 1000 * P) * 2 * np.pi
    a_expr, a_ent = draws[:1000 * P].reshape(1000, P), draws[1000 * P:].reshape(1000, P)
    F = orc.fidelity_samples(orc.run(specs, 4, a_expr))
    assert np.abs(F[:4096] - golden["c1/F_head"]).max() < ATOL
    assert abs(F.sum() - golden["c1/F_sum"]) < 1e-8
    prob, mid, counts = orc.gen_histo(F)
    assert len(counts) == 3746
    # counts may move only where an F sits within rounding of a bin edge
    assert np.abs(counts - golden["c1/hist"]).sum() <= 4
    assert abs(orc.expr(F, 16) - float(golden["c1/expr"])) < 1e-6
    assert abs(orc.expr_from_counts(golden["c1/hist"], 16) - float(golden["c1/expr"])) < 1e-12
    for N, ref in zip((4, 8.5, 16, 64), golden["c1/expr_altN"]):
        assert abs(orc.expr(F, N) - ref) < 1e-6 * max(1, abs(ref))
    assert abs(orc.expr(F, 16) - float(golden["c1/expr_filt"])) < 1e-6     # filt is a no-op
    ent = [orc.single_Q(s, 4) for s in orc.run(specs, 4, a_ent)]
    assert np.abs(np.array(ent) - golden["c1/ent"]).max() < ATOL


def test_efficient_measurements(golden):
    specs, _ = orc.generate_circuit("generic_HE", 4, 2)
    P = orc.n_params(specs)
    ang = (np.random.default_rng(1).random(40 * P) * 2 * np.pi).reshape(40, P)
    st = orc.run(specs, 4, ang)
    d = orc.efficient_measurements(st, 4)
    assert abs(d["Expr"] - float(golden["effm/expr"])) < 1e-6
    assert np.abs(np.array(d["Ent"]) - golden["effm/ent"]).max() < ATOL
    assert np.abs(np.array(d["Magic"]) - golden["effm/magic"]).max() < 1e-8
    assert np.abs(np.array(d["GKP"]) - golden["effm/gkp"]).max() < 1e-8
    f = orc.efficient_measurements(st, 4, full_data=True)
    assert np.abs(np.array(f["Expr"]) - golden["effm/full_expr"]).max() < ATOL
    assert np.abs(np.array(f["Ent"]) - golden["effm/full_ent"]).max() < ATOL
    assert np.abs(np.array(f["Magic"]) - golden["effm/full_magic"]).max() < 1e-8
    assert np.abs(np.array(f["GKP"]) - golden["effm/full_gkp"]).max() < 1e-8
    # 7 <= n < 12: overlaps taken, KL skipped -> -1 sentinel; skipped measures -> [-1,-1]
    specs7, _ = orc.generate_circuit("generic_HE", 7, 1)
    a7 = (np.random.default_rng(1).random(5 * orc.n_params(specs7)) * 2 * np.pi).reshape(5, -1)
    d7 = orc.efficient_measurements(orc.run(specs7, 7, a7), 7, measure_eom=False,
                                    measure_GKP=False)
    assert d7["Expr"] == -1 == float(golden["effm/n7_expr"])
    assert np.abs(np.array(d7["Ent"]) - golden["effm/n7_ent"]).max() < ATOL
    assert list(golden["effm/n7_magic"]) == [-1, -1] == d7["Magic"]
    # too few samples -> zero bins -> ValueError, as the reference (make_golden.py)
    with pytest.raises(ValueError):
        orc.efficient_measurements(st[:12], 4)


def test_example_script_values(golden):
    """/root/reference/example.py:18-20 on the module RNG stream."""
    specs, n, _ = oracle_case("example4")
    P = orc.n_params(specs)
    d = np.random.default_rng(1).random(300 * P) * 2 * np.pi
    F = orc.fidelity_samples(orc.run(specs, n, d[:150 * P].reshape(150, P)))
    assert abs(orc.expr(F, 16) - float(golden["example/expr150"])) < 1e-6
    st = orc.run(specs, n, d[150 * P:].reshape(150, P))
    assert abs(np.mean([orc.renyi(s) for s in st]) - float(golden["example/eom150"])) < 1e-8


def test_magic_12q(golden):
    """BASELINE config 4 shape, pinned to the reference's dense 4096^3 formulation."""
    specs, _ = orc.generate_circuit("NPQC", 12, 3)
    st = orc.run(specs, 12, golden["magic12/angles"])
    assert np.abs(st[0] - golden["magic12/state"]).max() < ATOL
    r2, g = float(golden["magic12/renyi2"]), float(golden["magic12/gkp"])
    assert abs(orc.renyi_fwht(st[0], 2.0) - r2) < RTOL * abs(r2)
    assert abs(orc.gkp(st[0]) - g) < RTOL * abs(g)
    assert abs(orc.single_Q(st[0], 12) - float(golden["magic12/Q"])) < ATOL


def test_noncommuting_shared_parameter_branch(golden_r3):
    """gates.py:458-466 with members that do not commute (and non-symmetric R_y members, so the
    element-wise conjugate of the block is not its inverse): the oracle's restatement against the
    reference-generated fixtures of tests/golden/make_golden_r3.py."""
    import cases_r3
    specs = [("R_y", 0), ("R_y", 1), ("R_y", 2), ("CHAIN", "CNOT"),
             ("shared_parameter", [("R_x", 0), ("R_z", 0)], False),
             ("shared_parameter", [("R_y", 0), ("R_y", 1)], False),
             ("CHAIN", "CPHASE"),
             ("shared_parameter", [("R_zz", 0, 1), ("R_x", 1), ("R_y", 2)], False),
             ("R_x", 2)]
    ang = np.array(cases_r3.NONCOMM3_ANGLES)
    st = orc.run(specs, 3, ang, None)
    gr = orc.gradients(specs, 3, ang, None)
    for k in range(len(ang)):
        assert np.abs(st[k] - golden_r3[f"noncomm3/{k}/state"]).max() < ATOL
        assert np.abs(gr[k] - golden_r3[f"noncomm3/{k}/grads"]).max() < ATOL
        F = orc.qfi(st[k], gr[k])
        ref = golden_r3[f"noncomm3/{k}/qfi"]
        assert np.abs(F - ref).max() < RTOL * max(1.0, np.abs(ref).max())
        assert orc.eqd(F, 1e-12) == int(golden_r3[f"noncomm3/{k}/eqd"])
        assert abs(orc.cost_zz(st[k:k + 1])[0] - golden_r3[f"noncomm3/{k}/cost"]) < ATOL
