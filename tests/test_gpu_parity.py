"""GPU parity: the CUDA engine (through the reference-facing API and the C ABI) against
the reference-generated goldens and the numpy oracle.  Tolerances are BASELINE.json's:
amplitudes / fidelities 1e-10 absolute, QFIM / magic 1e-8 relative, KL 1e-6."""
import os
import sys

import numpy as np
import pytest
import torch

import cases
import cases_r2
import pyramaterised_b200 as pyqc
from helpers import oracle_case, specs_from_circuit
from oracle import pqc_oracle as orc
from pyramaterised_b200 import engine

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ALL = sorted(cases.CASES)
ATOL = 1e-10
RTOL = 1e-8
GRAD_OK = [c for c in ALL if cases.CASES[c][2] > 0]


def reseed():
    pyqc.gates.rng.bit_generator.state = np.random.default_rng(1).bit_generator.state


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max())


@pytest.mark.parametrize("name", ALL)
def test_states_and_measures_vs_golden(golden, name):
    qc = cases.CASES[name][0](pyqc)
    ang = golden[f"{name}/angles"]
    m = pyqc.measure.Measurements(qc)
    st = qc.run_batch(ang)
    assert np.abs(st.cpu().numpy() - golden[f"{name}/states"]).max() < ATOL
    # single-sample API path (PQC.run / cost / single_Q), as the reference is called
    s0 = qc.run(list(ang[0]))
    assert np.abs(s0.numpy() - golden[f"{name}/states"][0]).max() < ATOL
    assert abs(qc.cost(list(ang[1])) - golden[f"{name}/cost"][1]) < ATOL
    assert abs(m.single_Q(s0, qc.n_qubits) - golden[f"{name}/Q"][0]) < ATOL
    assert np.abs(engine.meyer_wallach(st).cpu().numpy() - golden[f"{name}/Q"]).max() < ATOL
    _, F = engine.fidelity_hist(st, want_F=True)
    assert np.abs(F.cpu().numpy() - golden[f"{name}/F"]).max() < ATOL
    if cases.CASES[name][3]:
        mg = engine.magic(st, (2.0, 0.5)).cpu().numpy()
        assert rel(mg[0], golden[f"{name}/renyi2"]) < RTOL
        assert rel(mg[1] / (2 * np.log(2)), golden[f"{name}/gkp"]) < RTOL
        assert abs(m.renyi_entropy_fast(s0) - golden[f"{name}/renyi2"][0]) < RTOL
        assert abs(m.gkp_fast(s0) - golden[f"{name}/gkp"][0]) < RTOL


@pytest.mark.parametrize("name", GRAD_OK)
def test_gradients_qfi_eqd_vs_golden(golden, name):
    qc = cases.CASES[name][0](pyqc)
    G = cases.CASES[name][2]
    m = pyqc.measure.Measurements(qc)
    for s in range(G):
        a = list(golden[f"{name}/angles"][s])
        qc.update_state(a)
        grads = qc.get_gradients()
        got = np.stack([g.numpy() for g in grads])
        assert np.abs(got - golden[f"{name}/grads"][s]).max() < ATOL
        F = m.get_QFI(grad_list=grads)
        ref = golden[f"{name}/qfi"][s]
        assert rel(F, ref) < RTOL
        assert m.get_effective_quantum_dimension(1e-12) == int(golden[f"{name}/eqd"][s])
        assert abs(m.new_measure(F) - golden[f"{name}/new_measure"][s]) < 1e-8
        gv = m.get_gradient_vector(a)
        assert np.abs(np.array(gv) - golden[f"{name}/gradvec"][s]).max() < 1e-9
    # fused batch path agrees with the reference-shaped path
    Fb, eq = m.qfim_batch(golden[f"{name}/angles"][:G], cutoff_eigvals=1e-12)
    assert rel(Fb.cpu().numpy(), golden[f"{name}/qfi"]) < RTOL
    assert list(eq.cpu().numpy()) == list(golden[f"{name}/eqd"])


def test_quirk_q2_is_fenced_not_imitated():
    """fSim behind a non-parameterised gate: the reference would differentiate the wrong gate
    (circuit.py:186-189); the replacement refuses instead of silently diverging."""
    qc = pyqc.PQC(3)
    qc.add_layer([pyqc.H(0, 3), pyqc.fSim([0, 1], 3), pyqc.R_x(2, 3)])
    qc.update_state([0.3, 0.4, 0.5])
    with pytest.raises(NotImplementedError):
        qc.get_gradients()


def test_reference_known_answers(golden):
    """Ports of /root/reference/tests.py hard asserts."""
    from math import isclose
    c = pyqc.PQC(1)
    c.add_layer([pyqc.fixed_R_y(0, 1, np.pi / 2)])
    out = c.run("random")
    assert isclose(np.real(out[1][0][0]), 1 / np.sqrt(2))                 # tests.py:64-76
    c = pyqc.PQC(1)
    c.add_layer([pyqc.H(0, 1)], n=2)
    assert c.run("random") == pyqc.qt.basis(2, 0)                          # tests.py:81-86
    qg = cases.build_qg4(pyqc)
    assert isclose(qg.cost(cases.QG_ANGLES), cases.QG_ENERGY, abs_tol=1e-5)   # tests.py:207-209
    m = pyqc.measure.Measurements(qg)
    assert m.get_effective_quantum_dimension(10 ** -12) == cases.QG_EQD       # tests.py:211-212
    for N in (4, 6, 8):                                                       # tests.py:114-128
        for P in (1, 2, 2 ** (N // 2)):
            layers, th = pyqc.templates.NPQC_layers(P, N)
            npqc = pyqc.PQC(N)
            for l in layers:
                npqc.add_layer(l)
            npqc.state = npqc.run(angles=th)
            Q = pyqc.measure.Measurements(npqc).get_QFI()
            assert np.abs(Q - np.eye(len(Q))).max() < 1e-12
    bell = pyqc.State(golden["bell/state"])                                   # tests.py:284-295

    class Bell:
        n_qubits = 2
        state = bell

        def run(self, a):
            return bell

    bm = pyqc.measure.Measurements(Bell())
    assert isclose(bm.renyi_entropy_fast(bell), 0, abs_tol=1e-10)
    assert isclose(bm.gkp_fast(bell), 0, abs_tol=1e-10)
    assert isclose(bm.single_Q(bell, 2), 1, abs_tol=1e-10)


def test_config1_expressibility_entanglement(golden):
    """BASELINE config 1 through the reference API on the module RNG (quirk Q13)."""
    qc = pyqc.templates.generate_circuit("NPQC", 4, 4)
    m = pyqc.measure.Measurements(qc)
    reseed()
    e = m.expressibility(1000)
    ent = m.entanglement(1000)
    assert abs(e - float(golden["c1/expr"])) < 1e-6
    assert np.abs(np.array(ent) - golden["c1/ent"]).max() < ATOL
    reseed()
    F = np.array(m._gen_f_samples(1000))
    assert np.abs(F[:4096] - golden["c1/F_head"]).max() < ATOL
    assert abs(F.sum() - float(golden["c1/F_sum"])) < 1e-8
    # histogram kernel == np.histogram on the very same samples (bit exact integer counts)
    bins = engine.n_bins(len(F))
    assert bins == 3746
    ours = engine.hist_f64(torch.as_tensor(F, device="cuda"), bins).cpu().numpy()
    assert np.array_equal(ours, np.histogram(F, bins=bins, range=(0, 1))[0])
    assert np.abs(ours - golden["c1/hist"]).sum() <= 4
    for N, ref in zip((4, 8.5, 16, 64), golden["c1/expr_altN"]):
        assert abs(m.expr(list(F), N) - ref) < 1e-6 * max(1, abs(ref))
    prob, mid = m._gen_histo(list(F))
    assert abs(prob.sum() - 1) < 1e-12 and len(mid) == bins


def test_histogram_edge_semantics():
    """np.histogram(range=(0,1)): right-closed last bin, out-of-range dropped."""
    rng = np.random.default_rng(5)
    for bins in (1, 7, 83, 3746, 100003):
        edges = np.linspace(0, 1, bins + 1)
        F = np.concatenate([rng.random(5000), edges[:: max(1, bins // 500)],
                            np.nextafter(edges[1:-1:max(1, bins // 300)], 0),
                            np.nextafter(edges[1:-1:max(1, bins // 300)], 1),
                            [0.0, 1.0, 1.0 + 2e-16, -1e-18, 1.5, np.nan]])
        ours = engine.hist_f64(torch.as_tensor(F, device="cuda"), bins).cpu().numpy()
        assert np.array_equal(ours, np.histogram(F[~np.isnan(F)], bins=bins, range=(0, 1))[0]), bins
    with pytest.raises(ValueError):
        engine.hist_f64(torch.zeros(4, dtype=torch.float64, device="cuda"), 0)


def test_kl_haar_degenerate_cases_match_scipy():
    """expr() where the Haar weights (N-1)(1-F)^(N-2) underflow (config 5: N = 2^28): with every
    weight 0 the reference divides 0 / 0 and scipy.special.kl_div propagates NaN; at the full-size
    bin count the first bin survives and KL = 0; an empty PQC bin against a positive Haar weight
    contributes the weight."""
    for bins, total, N in ((3928, 523776, 2.0 ** 28), (374962, 49995000, 2.0 ** 28), (60, 1000, 16.0)):
        counts = np.zeros(bins, dtype=np.int64)
        counts[0] = total
        got = float(engine.kl_haar(torch.as_tensor(counts, device="cuda"), N).item())
        with np.errstate(all="ignore"):
            want = orc.expr_from_counts(counts, N)
        assert (np.isnan(want) and np.isnan(got)) or abs(got - want) <= 1e-12 * max(1.0, abs(want)), (bins, got, want)


def test_efficient_measurements(golden):
    qc = pyqc.templates.generate_circuit("generic_HE", 4, 2)
    m = pyqc.measure.Measurements(qc)
    reseed()
    d = m.efficient_measurements(40)
    assert abs(d["Expr"] - float(golden["effm/expr"])) < 1e-6
    assert np.abs(np.array(d["Ent"]) - golden["effm/ent"]).max() < ATOL
    assert np.abs(np.array(d["Magic"]) - golden["effm/magic"]).max() < 1e-8
    assert np.abs(np.array(d["GKP"]) - golden["effm/gkp"]).max() < 1e-8
    reseed()
    f = m.efficient_measurements(40, full_data=True)
    assert np.abs(np.array(f["Expr"]) - golden["effm/full_expr"]).max() < ATOL
    assert np.abs(np.array(f["Ent"]) - golden["effm/full_ent"]).max() < ATOL
    assert np.abs(np.array(f["Magic"]) - golden["effm/full_magic"]).max() < 1e-8
    assert np.abs(np.array(f["GKP"]) - golden["effm/full_gkp"]).max() < 1e-8
    import random
    random.seed(7)
    with pytest.raises(ValueError):
        m.efficient_measurements(12, angles="clifford")       # zero bins, as the reference
    random.seed(7)
    c = m.efficient_measurements(20, angles="clifford")
    assert np.abs(np.array(c["Magic"]) - golden["effm/cliff_magic"]).max() < 1e-8
    assert np.abs(np.array(c["Ent"]) - golden["effm/cliff_ent"]).max() < ATOL
    assert abs(c["Expr"] - float(golden["effm/cliff_expr"])) < 1e-6
    qc7 = pyqc.templates.generate_circuit("generic_HE", 7, 1)
    reseed()
    d7 = pyqc.measure.Measurements(qc7).efficient_measurements(5, measure_eom=False,
                                                                measure_GKP=False)
    assert d7["Expr"] == -1 and d7["Magic"] == [-1, -1] and d7["GKP"] == [-1, -1]
    assert np.abs(np.array(d7["Ent"]) - golden["effm/n7_ent"]).max() < ATOL
    z = m.efficient_measurements(0)
    assert z == {"Expr": -1, "Ent": [-1, -1], "Magic": [-1, -1], "GKP": [-1, -1]}


def test_example_script_values(golden):
    ex = cases.build_example4(pyqc)
    m = pyqc.measure.Measurements(ex)
    reseed()
    assert abs(m.expressibility(150) - float(golden["example/expr150"])) < 1e-6
    assert abs(m.entropy_of_magic(150) - float(golden["example/eom150"])) < 1e-8


def test_magic_12q_config4(golden):
    qc = pyqc.templates.generate_circuit("NPQC", 12, 3)
    st = qc.run_batch(golden["magic12/angles"])
    assert np.abs(st[0].cpu().numpy() - golden["magic12/state"]).max() < ATOL
    mg = engine.magic(st, (2.0, 0.5)).cpu().numpy()
    assert abs(mg[0, 0] - float(golden["magic12/renyi2"])) < RTOL * 4
    assert abs(mg[1, 0] / (2 * np.log(2)) - float(golden["magic12/gkp"])) < RTOL * 5
    assert abs(engine.meyer_wallach(st)[0].item() - float(golden["magic12/Q"])) < ATOL
    # a batch of random 12-qubit NPQC states against the FWHT oracle
    specs, _ = orc.generate_circuit("NPQC", 12, 3)
    ang = np.random.default_rng(3).random((3, orc.n_params(specs))) * 2 * np.pi
    st = qc.run_batch(ang)
    ref = orc.run(specs, 12, ang)
    assert np.abs(st.cpu().numpy() - ref).max() < ATOL
    mg = engine.magic(st, (2.0, 0.5)).cpu().numpy()
    for s in range(3):
        assert abs(mg[0, s] - orc.renyi_fwht(ref[s], 2.0)) < RTOL * 4
        assert abs(mg[1, s] - orc.renyi_fwht(ref[s], 0.5)) < RTOL * 10


def test_magic_register_blocked_kernel_matches_generic_and_oracle(monkeypatch):
    """n = 12 runs the register-blocked FWHT kernel (k_magic12); PQC_MAGIC=generic forces the
    in-memory FWHT used for every other n.  Both against the oracle's FWHT formulation of
    measure.py:318-349 on random and on structured states, several alphas, ragged mask slices."""
    rng = np.random.default_rng(12)
    st = rng.normal(size=(5, 4096)) + 1j * rng.normal(size=(5, 4096))
    st /= np.linalg.norm(st, axis=1, keepdims=True)
    st[3] = 0
    st[3, 77] = 1.0                                   # basis state: stabilizer, magic 0
    st[4] = 1 / 64.0                                  # |+>^12: stabilizer, magic 0
    dev = torch.as_tensor(st, device="cuda")
    alphas = (2.0, 0.5, 3.0)
    got = engine.magic(dev, alphas).cpu().numpy()
    monkeypatch.setenv("PQC_MAGIC", "generic")
    gen = engine.magic(dev, alphas).cpu().numpy()
    monkeypatch.delenv("PQC_MAGIC")
    assert np.abs(got - gen).max() < 1e-10
    for q, s in ((0, 0), (1, 1), (2, 2), (0, 3), (1, 4)):
        assert abs(got[q, s] - orc.renyi_fwht(st[s], alphas[q])) < 1e-9
    assert np.abs(got[:, 3:]).max() < 1e-10
    big = engine.magic(dev[:1].repeat(700, 1), (2.0,)).cpu().numpy()   # many CTAs per sample
    assert np.abs(big - got[0, 0]).max() < 1e-10


@pytest.mark.parametrize("kind,n,p", [("generic_HE", 13, 2), ("NPQC", 14, 3), ("TFIM", 16, 2),
                                      ("XXZ", 16, 1), ("qg_circuit", 15, 1), ("Circuit_2", 17, 1)])
def test_multi_pass_states_vs_oracle(kind, n, p):
    """n > tile bits: several HBM passes per circuit; amplitudes vs the numpy oracle."""
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    specs, init = orc.generate_circuit(kind, n, p)
    assert specs_from_circuit(qc) == specs
    ang = np.random.default_rng(n * 100 + p).random((2, orc.n_params(specs))) * 2 * np.pi
    got = qc.run_batch(ang).cpu().numpy()
    ref = orc.run(specs, n, ang, init)
    assert np.abs(got - ref).max() < ATOL
    Q = engine.meyer_wallach(torch.as_tensor(ref, device="cuda")).cpu().numpy()
    assert np.abs(Q - [orc.single_Q(r, n) for r in ref]).max() < ATOL


@pytest.mark.parametrize("kind,n,p", [("TFIM", 13, 2), ("XXZ", 14, 1), ("TFIM", 16, 1),
                                      ("generic_HE", 9, 2), ("NPQC", 10, 4), ("zfsim", 9, 2),
                                      ("fsim", 12, 1), ("fermionic", 8, 1), ("Circuit_9", 8, 2),
                                      ("TFIM_modified", 12, 2), ("qg_circuit", 11, 1)])
def test_multi_pass_qfim_vs_oracle(kind, n, p):
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    specs, init = orc.generate_circuit(kind, n, p)
    ang = np.random.default_rng(n + p).random((2, orc.n_params(specs))) * 2 * np.pi
    F, st = qc.qfim_batch(ang, want_states=True)
    ref_st = orc.run(specs, n, ang, init)
    assert np.abs(st.cpu().numpy() - ref_st).max() < ATOL
    gr = orc.gradients(specs, n, ang, init)
    for s in range(2):
        assert rel(F[s].cpu().numpy(), orc.qfi(ref_st[s], gr[s])) < RTOL


@pytest.mark.parametrize("kind,n,p", [("TFIM", 13, 8), ("XXZ", 12, 3), ("generic_HE", 12, 3),
                                      ("NPQC", 12, 4), ("Circuit_9", 11, 3)])
def test_meet_in_the_middle_qfim_matches_forward_plan_and_oracle(kind, n, p, monkeypatch):
    """The default QFIM plan differentiates the two halves of the circuit from both ends and takes
    the Gram matrix at the cut; PQC_BIDIR=0 runs the forward-only plan.  Both must agree with
    each other and with the oracle's literal measure.py:33-71."""
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    assert "BIDIR" in qc.program.describe()
    specs, init = orc.generate_circuit(kind, n, p)
    ang = np.random.default_rng(7 * n + p).random((3, orc.n_params(specs))) * 2 * np.pi
    F, st = qc.qfim_batch(ang, want_states=True)
    monkeypatch.setenv("PQC_BIDIR", "0")
    F0, st0 = qc.qfim_batch(ang, want_states=True)
    monkeypatch.delenv("PQC_BIDIR")
    assert np.abs((st - st0).cpu().numpy()).max() < ATOL
    assert rel(F.cpu().numpy(), F0.cpu().numpy()) < 1e-12
    ref_st = orc.run(specs, n, ang[:1], init)
    gr = orc.gradients(specs, n, ang[:1], init)
    assert rel(F[0].cpu().numpy(), orc.qfi(ref_st[0], gr[0])) < RTOL


@pytest.mark.parametrize("kind,n,p", [("TFIM", 16, 3), ("TFIM", 13, 4), ("TFIM_modified", 14, 2)])
def test_layer_pass_fast_path_matches_generic_sweep_kernel(kind, n, p, monkeypatch):
    """Passes made of aligned nibble sweeps run on k_layer_pass (plan lines say fast=1);
    PQC_FAST=0 forces the generic k_sweep_pass.  Same arithmetic in the same order: states and
    QFIMs must agree to rounding, and with the oracle."""
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    assert "fast=1" in qc.program.describe()
    specs, init = orc.generate_circuit(kind, n, p)
    ang = np.random.default_rng(n + 31 * p).random((3, orc.n_params(specs))) * 2 * np.pi
    st = qc.run_batch(ang)
    F = qc.qfim_batch(ang)
    monkeypatch.setenv("PQC_FAST", "0")
    st0 = qc.run_batch(ang)
    F0 = qc.qfim_batch(ang)
    monkeypatch.delenv("PQC_FAST")
    assert np.abs((st - st0).cpu().numpy()).max() < 1e-14
    assert rel(F.cpu().numpy(), F0.cpu().numpy()) < 1e-12
    ref = orc.run(specs, n, ang[:1], init)
    assert np.abs(st[:1].cpu().numpy() - ref).max() < ATOL
    gr = orc.gradients(specs, n, ang[:1], init)
    assert rel(F[0].cpu().numpy(), orc.qfi(ref[0], gr[0])) < RTOL


@pytest.mark.parametrize("kind,n,p", [("XXZ", 16, 2), ("XXZ", 13, 3), ("XXZ", 12, 4),
                                      ("NPQC", 14, 3), ("NPQC", 12, 5), ("NPQC", 17, 2)])
def test_layer_sequence_path_matches_generic_sweep_kernel(kind, n, p, monkeypatch):
    """XXZ passes (XY pair rotations, nibble sweeps in any order) and NPQC passes (runs of R_z
    phases and CZ signs, applied once per run) execute on k_layer_seq (plan lines say fast=2);
    PQC_SEQ=0 keeps them on the generic k_sweep_pass.  States, derivative states and QFIMs agree
    to rounding, and row 0 agrees with the oracle."""
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    assert "fast=2" in qc.program.describe()
    specs, init = orc.generate_circuit(kind, n, p)
    ang = np.random.default_rng(3 * n + p).random((3, orc.n_params(specs))) * 2 * np.pi
    st = qc.run_batch(ang)
    F = qc.qfim_batch(ang)
    gr = qc.program.gradients(ang, init=qc.initial_state.tensor)
    monkeypatch.setenv("PQC_SEQ", "0")
    st0 = qc.run_batch(ang)
    F0 = qc.qfim_batch(ang)
    gr0 = qc.program.gradients(ang, init=qc.initial_state.tensor)
    monkeypatch.delenv("PQC_SEQ")
    assert np.abs((st - st0).cpu().numpy()).max() < 1e-14
    assert np.abs((gr - gr0).cpu().numpy()).max() < 1e-13
    assert rel(F.cpu().numpy(), F0.cpu().numpy()) < 1e-12
    ref = orc.run(specs, n, ang[:1], init)
    assert np.abs(st[:1].cpu().numpy() - ref).max() < ATOL
    g1 = orc.gradients(specs, n, ang[:1], init)
    assert np.abs(gr[:1, 1:].cpu().numpy() - g1).max() < ATOL
    assert rel(F[0].cpu().numpy(), orc.qfi(ref[0], g1[0])) < RTOL


@pytest.mark.parametrize("kind,n,p,S", [("TFIM", 16, 3, 37), ("TFIM", 13, 4, 5), ("TFIM_modified", 14, 2, 3),
                                        ("XXZ", 16, 2, 19), ("XXZ", 13, 3, 3), ("NPQC", 14, 3, 3),
                                        ("NPQC", 17, 2, 3), ("TFIM", 20, 1, 2)])
def test_tile_pipe_kernel_matches_per_tile_kernels_bitwise(kind, n, p, S, monkeypatch):
    """k_tile_pipe (persistent CTAs, TMA-fed ring of tiles, runtime sweep geometry) runs the same
    plans (PQC_PIPE=1) with the same arithmetic in the same order as k_layer_pass / k_layer_seq:
    states and derivative states must be BIT-identical, QFIMs too; row 0 agrees with the oracle."""
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    specs, init = orc.generate_circuit(kind, n, p)
    ang = np.random.default_rng(5 * n + p).random((S, orc.n_params(specs))) * 2 * np.pi
    l0 = engine.launch_count()
    monkeypatch.setenv("PQC_PIPE", "1")
    st = qc.run_batch(ang)
    F = qc.qfim_batch(ang)
    gr = qc.program.gradients(ang[:3], init=qc.initial_state.tensor)
    monkeypatch.setenv("PQC_PIPE", "0")
    st0 = qc.run_batch(ang)
    F0 = qc.qfim_batch(ang)
    gr0 = qc.program.gradients(ang[:3], init=qc.initial_state.tensor)
    monkeypatch.delenv("PQC_PIPE")
    assert engine.launch_count() > l0
    assert torch.equal(torch.view_as_real(st), torch.view_as_real(st0))
    assert torch.equal(torch.view_as_real(gr), torch.view_as_real(gr0))
    assert torch.equal(F, F0)
    ref = orc.run(specs, n, ang[:1], init)
    assert np.abs(st[:1].cpu().numpy() - ref).max() < ATOL
    if n <= 16:
        g1 = orc.gradients(specs, n, ang[:1], init)
        assert rel(F[0].cpu().numpy(), orc.qfi(ref[0], g1[0])) < RTOL


FRONT_CASES = [("XXZ", 16, 3, 5), ("XXZ", 13, 4, 3), ("generic_HE", 16, 3, 5), ("generic_HE", 12, 4, 3),
               ("NPQC", 16, 5, 5), ("NPQC", 17, 3, 2), ("TFIM", 16, 3, 5), ("TFIM_modified", 14, 2, 3),
               ("qg_circuit", 13, 2, 3), ("Circuit_2", 12, 2, 3), ("clifford", 12, 2, 3),
               ("Circuit_9", 12, 2, 3), ("y_CPHASE", 13, 2, 3), ("XXZ", 20, 2, 1), ("XXZ", 16, 16, 2),
               ("XXZ", 14, 6, 3), ("TFIM", 14, 5, 2)]


@pytest.mark.parametrize("kind,n,p,S", FRONT_CASES)
def test_front_plan_matches_block_plan_and_oracle(kind, n, p, S, monkeypatch):
    """PQC.run through the front planner (light-cone passes on k_tile_pipe: any 4 tile bits in
    registers, X / CNOT folded into the sweeps' addresses, R_z in tangent form, per-level
    table-lookup R_zz phases; PQC_FRONT=1) against the block planner's plan on the per-tile
    kernels (PQC_FRONT=0) and against the oracle -- the three-way check of every new path."""
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    assert "FRONT plan" in qc.program.describe()
    specs = specs_from_circuit(qc)
    init = qc.initial_state.tensor.cpu().numpy()
    ang = np.random.default_rng(7 * n + p).random((S, max(1, orc.n_params(specs)))) * 2 * np.pi
    l0 = engine.launch_count()
    monkeypatch.setenv("PQC_FRONT", "1")
    st = qc.run_batch(ang)
    monkeypatch.setenv("PQC_FRONT", "0")
    st0 = qc.run_batch(ang)
    monkeypatch.delenv("PQC_FRONT")
    assert engine.launch_count() > l0
    assert np.abs((st - st0).cpu().numpy()).max() < 1e-12
    ref = orc.run(specs, n, ang[:1], init)
    assert np.abs(st[:1].cpu().numpy() - ref).max() < ATOL
    nrm = engine.overlap(st, st).cpu().numpy()
    assert np.abs(nrm - 1).max() < 1e-12


@pytest.mark.parametrize("kind,n,p", [("XXZ", 16, 16), ("generic_HE", 16, 16), ("NPQC", 16, 16),
                                      ("NPQC", 20, 20), ("NPQC", 22, 20), ("generic_HE", 20, 8),
                                      ("XXZ", 20, 8), ("TFIM", 20, 8), ("qg_circuit", 18, 4)])
def test_full_depth_states_vs_oracle(kind, n, p):
    """Full-depth circuits of BASELINE configs 3 / 5 on the default plan (front planner) against the
    oracle: row 0 of a small batch, every amplitude to 1e-10."""
    qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
    specs = specs_from_circuit(qc)
    init = qc.initial_state.tensor.cpu().numpy()
    ang = np.random.default_rng(n + p).random((2, orc.n_params(specs))) * 2 * np.pi
    st = qc.run_batch(ang)
    ref = orc.run(specs, n, ang[:1], init)
    assert np.abs(st[0].cpu().numpy() - ref[0]).max() < ATOL
    assert abs(float(engine.meyer_wallach(st[:1])[0].item()) - orc.single_Q(ref[0], n)) < ATOL


@pytest.mark.parametrize("n,S", [(3, 7), (4, 1000), (7, 5), (10, 300), (11, 3), (12, 9), (13, 2),
                                 (14, 3), (15, 3), (16, 5), (19, 2), (20, 1), (23, 1), (24, 1)])
def test_meyer_wallach_tile_kernel_matches_generic_and_oracle(n, S, monkeypatch):
    """n >= 12: k_mw_tiles (all 12 tile bits per read, 8 more per further pass), n <= 11:
    k_mw_small (one CTA per state, one read) -- both with fixed-order reductions -- against
    k_mw_accumulate (PQC_MW=generic), the oracle, ptrace and themselves run twice (bitwise
    reproducible: no floating-point atomics)."""
    qc = pyqc.templates.generate_circuit("generic_HE", n, 2)
    ang = np.random.default_rng(n).random((S, qc.n_true_params)) * 2 * np.pi
    st = qc.run_batch(ang)
    Q1 = engine.meyer_wallach(st)
    Q2 = engine.meyer_wallach(st)
    assert torch.equal(Q1, Q2)
    monkeypatch.setenv("PQC_MW", "generic")
    Q0 = engine.meyer_wallach(st)
    rho0 = engine.ptrace_1q(st[0], n - 3).cpu().numpy()
    monkeypatch.delenv("PQC_MW")
    rho1 = engine.ptrace_1q(st[0], n - 3).cpu().numpy()
    assert np.abs((Q1 - Q0).cpu().numpy()).max() < 1e-12
    assert np.abs(rho1 - rho0).max() < 1e-12
    if n <= 20:
        assert abs(float(Q1[0].item()) - orc.single_Q(st[0].cpu().numpy(), n)) < ATOL
        assert abs(float(Q1[S - 1].item()) - orc.single_Q(st[S - 1].cpu().numpy(), n)) < ATOL


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11, 12, 13])
def test_meyer_wallach_on_random_states_every_small_size(n):
    """Haar-like random states (not circuit outputs: every amplitude generic, unnormalised rows too)
    at the sizes where the kernels change hands (k_mw_small up to 11 qubits, the tile kernel from
    12): Q and every reduced density matrix against the oracle (measure.py:226-237)."""
    g = np.random.default_rng(100 + n)
    S = 37
    st = g.normal(size=(S, 2 ** n)) + 1j * g.normal(size=(S, 2 ** n))
    st /= np.linalg.norm(st, axis=1, keepdims=True)
    d = torch.as_tensor(st, device="cuda")
    Q = engine.meyer_wallach(d).cpu().numpy()
    ref = np.array([orc.single_Q(s, n) for s in st])
    assert np.abs(Q - ref).max() < 1e-12
    for q in (0, n - 1):
        m = st[5].reshape(2 ** q, 2, 2 ** (n - q - 1)).transpose(1, 0, 2).reshape(2, -1)
        rho = engine.ptrace_1q(d[5], q).cpu().numpy().reshape(2, 2)
        assert np.abs(rho - m @ m.conj().T).max() < 1e-12


def test_ragged_and_empty_batches():
    """Batch edges: a QFIM batch that does not fill its last 256-set chunk, a single row, an
    empty batch; every row must equal the one-row call bit for bit (fixed summation orders)."""
    qc = pyqc.templates.generate_circuit("TFIM", 12, 3, shuffle=False)
    P = qc.n_true_params
    ang = np.random.default_rng(9).random((300, P)) * 2 * np.pi
    F = qc.qfim_batch(ang)
    st = qc.run_batch(ang)
    assert F.shape == (300, P, P) and st.shape == (300, 4096)
    for row in (0, 255, 256, 299):
        F1 = qc.qfim_batch(ang[row:row + 1])
        assert torch.equal(F1[0], F[row])
        assert torch.equal(qc.run_batch(ang[row:row + 1])[0], st[row])
    assert torch.equal(F, F.transpose(1, 2))                       # exactly symmetric
    e = np.zeros((0, P))
    assert qc.run_batch(e).shape == (0, 4096)
    assert qc.qfim_batch(e).shape == (0, P, P)
    assert engine.meyer_wallach(st[:0]).shape == (0,)
    assert engine.magic(st[:0], (2.0,)).shape == (1, 0)
    with pytest.raises(IndexError):
        qc.run_batch(ang[:, :P - 1])                               # short angle rows, circuit.py:97


def test_eigvalsh_vs_lapack():
    rng = np.random.default_rng(11)
    for P in (1, 2, 5, 12, 32, 33, 64):
        A = rng.normal(size=(6, P, P))
        A = A + A.transpose(0, 2, 1)
        A[1] = A[1] @ A[1].T                                   # PSD
        v = rng.normal(size=(P, max(1, P // 3)))
        A[2] = v @ v.T                                         # rank deficient (zeros)
        w = engine.eigvalsh(torch.as_tensor(A, device="cuda")).cpu().numpy()
        ref = np.linalg.eigvalsh(A)
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(w - ref) / scale).max() < 1e-13, P
        ww, vv = engine.eigh(torch.as_tensor(A, device="cuda"))
        ww, vv = ww.cpu().numpy(), vv.cpu().numpy()
        for k in range(6):
            assert np.abs(A[k] @ vv[k] - vv[k] * ww[k]).max() < 1e-11 * max(1, scale[k, 0])
            assert np.abs(vv[k].T @ vv[k] - np.eye(P)).max() < 1e-12
        cnt = engine.count_greater(torch.as_tensor(ref, device="cuda"), 1e-12).cpu().numpy()
        assert np.array_equal(cnt, (ref > 1e-12).sum(axis=1))


def test_size_independent_properties_16q():
    """At BASELINE config 3 size: unit norm, QFIM symmetric PSD, gauge invariance of F,
    and F(theta) of TFIM bounded by 4 Var <= (2 * #generators)^2."""
    qc = pyqc.templates.generate_circuit("TFIM", 16, 16)
    ang = np.random.default_rng(1).random((3, 32)) * 2 * np.pi
    F, st = qc.qfim_batch(ang, want_states=True)
    nrm = engine.overlap(st, st).cpu().numpy()
    assert np.abs(nrm - 1).max() < 1e-12
    Fn = F.cpu().numpy()
    assert np.abs(Fn - Fn.transpose(0, 2, 1)).max() == 0
    w = engine.eigvalsh(F).cpu().numpy()
    assert w.min() > -1e-9 and np.diagonal(Fn, axis1=1, axis2=2).max() <= 4 * 8 ** 2 + 1e-9
    eq = engine.count_greater(torch.as_tensor(w, device="cuda"), 1e-12).cpu().numpy()
    assert np.array_equal(eq, (np.linalg.eigvalsh(Fn) > 1e-12).sum(axis=1))


def test_eqd_stable_cutoff_agrees_across_plans_and_solvers():
    """BASELINE config 3's output (EQD, measure.py:77-87) on 256 parameter sets of the bench
    workload, TFIM 16q x 16 layers: at the cutoff 1e-10 -- two decades above the rounding noise of
    the rank-16 QFIM's 17th eigenvalue (profiles/r2_eqd_noise.json) -- the meet-in-the-middle
    plan, the forward-only plan and LAPACK on the same QFIM all give the same EQD for every set;
    at the reference's own 1e-12 (tests.py:211) the Jacobi kernel still equals LAPACK set for set
    on the same matrix (the cutoff then splits noise, whichever plan made the matrix)."""
    import os
    qc = pyqc.templates.generate_circuit("TFIM", 16, 16)
    ang = torch.from_numpy(np.random.default_rng(1).random((10000, 32))[:256] * 2 * np.pi).cuda()
    F = qc.qfim_batch(ang)
    w = engine.eigvalsh(F)
    wl = np.linalg.eigvalsh(F.cpu().numpy())
    os.environ["PQC_BIDIR"] = "0"
    try:
        qc2 = pyqc.templates.generate_circuit("TFIM", 16, 16)     # planned forward-only
        F2 = qc2.qfim_batch(ang)
    finally:
        del os.environ["PQC_BIDIR"]
    assert float((F - F2).abs().max() / F.abs().max()) < 1e-11
    w2 = engine.eigvalsh(F2)
    e10 = engine.count_greater(w, 1e-10).cpu().numpy()
    assert np.array_equal(e10, np.full(256, 16))
    assert np.array_equal(e10, engine.count_greater(w2, 1e-10).cpu().numpy())
    assert np.array_equal(e10, (wl > 1e-10).sum(axis=1))
    assert np.array_equal(engine.count_greater(w, 1e-12).cpu().numpy(), (wl > 1e-12).sum(axis=1))


def test_operator_algebra_and_state_surface(golden):
    """`Gate * state`, prod(), conj(), overlap, ptrace, == (gates.py:30-31,63-85)."""
    N = 3
    psi = pyqc.qt.tensor([pyqc.qt.basis(2, 0)] * N)
    g = pyqc.R_y(1, N)
    g.set_theta(0.7)
    chain = pyqc.CHAIN(pyqc.CNOT, N)
    out = chain * (g * psi)
    ref = orc.run([("R_y", 1), ("CHAIN", "CNOT")], N, [[0.7]])[0]
    assert np.abs(out.numpy() - ref).max() < ATOL
    out2 = pyqc.prod([g, chain][::-1]) * psi
    assert out2 == out
    U = g.operation.full()
    assert np.abs(U - np.kron(np.kron(np.eye(2), orc.rot_matrix("y", 0.7)), np.eye(2))).max() < 1e-15
    rx = pyqc.R_x(0, N)
    rx.set_theta(1.3)
    assert np.abs(rx.operation.conj().full() - rx.operation.full().conj()).max() < 1e-15
    assert abs(out.overlap(out) - 1) < 1e-14
    rho = out.ptrace(1)
    assert abs((rho * rho).tr() - np.trace(rho.full() @ rho.full()).real) < 1e-15
    assert out.dims == [[2] * N, [1] * N] and out.data.toarray().shape == (8, 1)
    d = g.derivative() * out
    assert np.abs(d.numpy() - orc.apply_1q(out.numpy()[None], N, 1, -0.5j * orc.SY)[0]).max() < ATOL


def test_dense_operators_act_on_states():
    """`ARBGATE * state` (gates.py:63-67 with the dense operation of gates.py:416-420) and the dense
    fSim derivative operators of gates.py:609-648, 719-737 acting on device kets; fsim_gate on a
    register as a symbolic operator whose dense form is the embedded 4 x 4 matrix."""
    import scipy.linalg
    G = pyqc.gates
    N, th, ph = 3, 0.37, 1.21
    rng = np.random.default_rng(11)
    v = rng.standard_normal(1 << N) + 1j * rng.standard_normal(1 << N)
    v /= np.linalg.norm(v)
    psi = pyqc.State(v)
    H = pyqc.templates.TFIM_hamiltonian(N, 0.8, 0.1)
    g = pyqc.ARBGATE(H)
    g.set_theta(0.6)
    out = g * psi
    assert out.dims == psi.dims
    assert np.abs(out.numpy().reshape(-1) - scipy.linalg.expm(-0.6j * H.full()) @ v).max() < ATOL
    for d, fn in ((1, G.fsim_gate_d_theta), (2, G.fsim_gate_d_phi)):
        got = fn(th, ph, N=N, control=2, target=0) * psi
        want = orc.apply_2q(v[None].copy(), N, 2, 0, orc.fsim_matrix(th, ph, d))[0]
        assert np.abs(got.numpy().reshape(-1) - want).max() < ATOL
    U = G.fsim_gate(th, ph, N=N, control=0, target=2).full()
    assert np.abs(U - G._expand_2toN(orc.fsim_matrix(th, ph, 0), N, 0, 2)).max() < 1e-14
    U = G.fixed_fsim_gate(th, N=N, control=1, target=2).full()
    assert np.abs(U - G._expand_2toN(orc.fixed_fsim_matrix(th, 0), N, 1, 2)).max() < 1e-14
    f = pyqc.fSim([0, 2], N)
    f.set_theta(th)
    f.set_phi(ph)
    assert np.abs((f * psi).numpy().reshape(-1)
                  - orc.apply_2q(v[None].copy(), N, 0, 2, orc.fsim_matrix(th, ph, 0))[0]).max() < ATOL


def test_fidelity_dmma_matches_fma_and_numpy(monkeypatch):
    """The FP64 tensor-core pair-fidelity kernel against the CUDA-core one and numpy, on
    ragged sizes, rectangular and triangular blocks, with identical integer histograms."""
    rng = np.random.default_rng(8)
    for n, SA, SB in ((4, 5, 3), (6, 70, 70), (9, 131, 64), (10, 200, 77)):
        D = 2 ** n
        A = rng.normal(size=(SA, D)) + 1j * rng.normal(size=(SA, D))
        A /= np.linalg.norm(A, axis=1, keepdims=True)
        B = rng.normal(size=(SB, D)) + 1j * rng.normal(size=(SB, D))
        B /= np.linalg.norm(B, axis=1, keepdims=True)
        tA, tB = torch.as_tensor(A, device="cuda"), torch.as_tensor(B, device="cuda")
        bins = 97
        monkeypatch.setenv("PQC_FIDELITY", "fma")
        h0, F0 = engine.fidelity_hist(tA, tB, bins=bins, want_F=True)
        t0, T0 = engine.fidelity_hist(tA, bins=bins, want_F=True)
        monkeypatch.delenv("PQC_FIDELITY")
        h1, F1 = engine.fidelity_hist(tA, tB, bins=bins, want_F=True)
        t1, T1 = engine.fidelity_hist(tA, bins=bins, want_F=True)
        ref = np.abs(A.conj() @ B.T) ** 2
        assert np.abs(F1.cpu().numpy() - ref).max() < 1e-13
        assert np.abs(F0.cpu().numpy() - ref).max() < 1e-13
        tri = np.abs(A.conj() @ A.T)[np.triu_indices(SA, 1)] ** 2
        assert np.abs(T1.cpu().numpy() - tri).max() < 1e-13
        assert np.abs(T0.cpu().numpy() - tri).max() < 1e-13
        assert int(h1.sum()) == SA * SB and int(t1.sum()) == SA * (SA - 1) // 2
        assert np.array_equal(h1.cpu().numpy(),
                              np.histogram(F1.cpu().numpy().ravel(), bins=bins, range=(0, 1))[0])
        assert np.abs(h1.cpu().numpy() - h0.cpu().numpy()).sum() <= 2


def test_fidelity_split_k_path(monkeypatch):
    """A few pairs of very long vectors (config 5) take the split-K path; forced here at n = 13
    and compared with numpy and with the tensor-core path (same integer histogram)."""
    rng = np.random.default_rng(5)
    D = 1 << 13
    A = rng.normal(size=(5, D)) + 1j * rng.normal(size=(5, D))
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    A[1] = A[0]                                          # fidelity exactly 1: last bin
    B = rng.normal(size=(3, D)) + 1j * rng.normal(size=(3, D))
    B /= np.linalg.norm(B, axis=1, keepdims=True)
    tA, tB = torch.as_tensor(A, device="cuda"), torch.as_tensor(B, device="cuda")
    h0, F0 = engine.fidelity_hist(tA, tB, bins=11, want_F=True)
    t0, T0 = engine.fidelity_hist(tA, bins=11, want_F=True)
    monkeypatch.setenv("PQC_FIDELITY", "splitk")
    h1, F1 = engine.fidelity_hist(tA, tB, bins=11, want_F=True)
    t1, T1 = engine.fidelity_hist(tA, bins=11, want_F=True)
    monkeypatch.delenv("PQC_FIDELITY")
    assert np.abs(F1.cpu().numpy() - np.abs(A.conj() @ B.T) ** 2).max() < 1e-13
    tri = np.abs(A.conj() @ A.T)[np.triu_indices(5, 1)] ** 2
    assert np.abs(T1.cpu().numpy() - tri).max() < 1e-13
    assert np.abs((F1 - F0).cpu().numpy()).max() < 1e-13
    assert torch.equal(h1, h0) and int(h1.sum()) == 15
    assert int(t1.sum()) == 10 and int(t1[-1]) >= 1 and np.abs((t1 - t0).cpu().numpy()).sum() <= 2


def test_fidelity_blocked_split_k_kernel(monkeypatch):
    """The blocked split-K kernel (each state read once per 32 x 4 pair tile; the form config 5's
    resident-block x travelling-block products take): ragged tiles on both sides, the triangular
    form, fidelity 1 in the last bin; same values as numpy and the same integer histogram as the
    tensor-core path; two runs are bitwise equal (fixed-order reductions)."""
    rng = np.random.default_rng(6)
    D = 1 << 13
    A = rng.normal(size=(37, D)) + 1j * rng.normal(size=(37, D))
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    A[9] = A[3]
    B = rng.normal(size=(6, D)) + 1j * rng.normal(size=(6, D))
    B /= np.linalg.norm(B, axis=1, keepdims=True)
    tA, tB = torch.as_tensor(A, device="cuda"), torch.as_tensor(B, device="cuda")
    h0, F0 = engine.fidelity_hist(tA, tB, bins=13, want_F=True)
    t0, T0 = engine.fidelity_hist(tA, bins=13, want_F=True)
    monkeypatch.setenv("PQC_FIDELITY", "block")
    h1, F1 = engine.fidelity_hist(tA, tB, bins=13, want_F=True)
    h2, F2 = engine.fidelity_hist(tA, tB, bins=13, want_F=True)
    t1, T1 = engine.fidelity_hist(tA, bins=13, want_F=True)
    monkeypatch.delenv("PQC_FIDELITY")
    assert torch.equal(F1, F2) and torch.equal(h1, h2)
    assert np.abs(F1.cpu().numpy() - np.abs(A.conj() @ B.T) ** 2).max() < 1e-13
    tri = np.abs(A.conj() @ A.T)[np.triu_indices(37, 1)] ** 2
    assert np.abs(T1.cpu().numpy() - tri).max() < 1e-13
    assert torch.equal(h1, h0) and int(h1.sum()) == 37 * 6
    assert int(t1.sum()) == 37 * 36 // 2 and int(t1[-1]) >= 1
    assert np.abs((t1 - t0).cpu().numpy()).sum() <= 2


# ---- size-independent properties at BASELINE.json's full sizes ------------------------------
def test_config2_full_size_properties():
    """generic_HE 10q x 10 layers, S = 1e5: 4 999 950 000 pairs into 37 499 625 bins."""
    S = 100000
    qc = pyqc.templates.generate_circuit("generic_HE", 10, 10)
    ang = torch.from_numpy(np.random.default_rng(1).random((S, qc.n_true_params)) * 2 * np.pi)
    st = qc.run_batch(ang.cuda())
    nrm = engine.overlap(st[:4096], st[:4096]).cpu().numpy()
    assert np.abs(nrm - 1).max() < 1e-12
    Q = engine.meyer_wallach(st).cpu().numpy()
    assert Q.min() > 0.5 and Q.max() <= 1 + 1e-12
    pairs = S * (S - 1) // 2
    bins = engine.n_bins(pairs)
    assert (pairs, bins) == (4999950000, 37499625)
    hist, _ = engine.fidelity_hist(st, bins=bins)
    assert int(hist.sum().item()) == pairs                       # every pair binned exactly once
    kl = float(engine.kl_haar(hist, 2.0 ** 10).item())
    assert 0 <= kl < 1e-2                                        # deep HE circuit ~ Haar
    # integer histogram is invariant under a permutation of the sample set (subset for time)
    sub = st[:3000]
    perm = torch.randperm(3000, generator=torch.Generator().manual_seed(3)).cuda()
    b = engine.n_bins(3000 * 2999 // 2)
    h1, _ = engine.fidelity_hist(sub, bins=b)
    h2, _ = engine.fidelity_hist(sub[perm].contiguous(), bins=b)
    assert torch.equal(h1, h2)
    # sharded form (dist.py partition, single rank) gives the same integer counts
    from pyramaterised_b200 import dist as pdist
    assert torch.equal(pdist.sharded_fidelity_hist(sub, b), h1)


def _sparse_circuit(n, q):
    """Three layers of every primitive family on the 12 qubits q[0..11] of an n-qubit register."""
    G = pyqc.gates
    qc = pyqc.PQC(n)
    qc.add_layer([G.H(q[0], n), G.fixed_R_y(q[3], n, 0.7)] + [G.R_y(k, n) for k in q] +
                 [G.CNOT([q[0], q[11]], n), G.CZ([q[5], q[6]], n), G.CNOT([q[7], q[2]], n)])
    qc.add_layer([G.R_z(k, n) for k in q] +
                 [G.R_xx([q[2], q[9]], n), G.R_zz([q[4], q[7]], n), G.R_yy([q[1], q[10]], n),
                  G.CNOT([q[8], q[3]], n)])
    qc.add_layer([G.R_x(k, n) for k in q] + [G.CZ([q[0], q[6]], n), G.R_zz([q[11], q[5]], n)])
    return qc


def test_config5_size_28_qubit_register_vs_12_qubit_oracle():
    """Config 5's register size (28 qubits, 4 GiB per complex128 state).  Gates act on 12
    qubits spread over the whole register (tile-low, middle and top index bits), so the state is
    the 12-qubit oracle state embedded at the other qubits = 0: amplitudes bit-for-bit
    comparable at full size, Meyer-Wallach Q_28 = 12/28 * Q_12, pair fidelity = the 12-qubit one."""
    n = 28
    q = [0, 1, 2, 9, 10, 13, 14, 20, 24, 25, 26, 27]
    big, small = _sparse_circuit(n, q), _sparse_circuit(12, list(range(12)))
    P = small.n_true_params
    assert big.n_true_params == P
    ang = np.random.default_rng(28).random((2, P)) * 2 * np.pi
    ref = orc.run(specs_from_circuit(small), 12, ang, None)
    st = big.run_batch(ang)                                     # [2, 2^28]
    assert st.shape == (2, 1 << n)
    # index of the embedded amplitude: bit (11 - k) of j -> index bit (n - 1 - q[k])
    j = np.arange(4096)
    idx = np.zeros(4096, dtype=np.int64)
    for k in range(12):
        idx |= ((j >> (11 - k)) & 1) << (n - 1 - q[k])
    got = st[:, torch.as_tensor(idx, device="cuda")].cpu().numpy()
    assert np.abs(got - ref).max() < ATOL
    nrm = engine.overlap(st, st).cpu().numpy()
    assert np.abs(nrm - 1).max() < 1e-12                        # nothing leaked elsewhere
    Q = engine.meyer_wallach(st).cpu().numpy()
    for s in range(2):
        assert abs(Q[s] - 12.0 / 28.0 * orc.single_Q(ref[s], 12)) < ATOL
    _, F = engine.fidelity_hist(st, bins=7, want_F=True)
    assert abs(float(F.reshape(-1)[0]) - abs(np.vdot(ref[0], ref[1])) ** 2) < ATOL


def test_streamed_expressibility_equals_resident_form():
    """The block-streamed expressibility (two blocks resident, blocks regenerated; the form
    config 5 needs) bins every pair exactly once: same KL and Q list as the resident form."""
    qc = pyqc.templates.generate_circuit("generic_HE", 10, 4)
    m = pyqc.measure.Measurements(qc)
    reseed()
    e0 = m.expressibility(333)
    reseed()
    q0 = m.entanglement(333)
    for block in (50, 333, 1000):
        reseed()
        e1, q1 = m.expressibility_streamed(333, block, want_Q=True)
        assert e1 == e0
        assert q1 == q0
    reseed()
    assert m.expressibility_streamed(1, 8) == 0


def test_config4_full_size_stabilizer_states_have_zero_magic():
    """NPQC 12q, S = 1000: with Clifford angles every state is a stabilizer state, so the
    Renyi-2 magic and the GKP magic vanish (cf. tests.py:284-295 for the Bell state), and for
    random angles 0 <= M2 <= ln((2^n + 1) / 2)."""
    qc = pyqc.templates.generate_circuit("NPQC", 12, 6)
    P = qc.n_true_params
    rng = np.random.default_rng(12)
    cl = rng.integers(0, 4, size=(1000, P)) * (np.pi / 2)
    mg = engine.magic(qc.run_batch(cl), (2.0, 0.5)).cpu().numpy()
    assert np.abs(mg[0]).max() < 1e-9 and np.abs(mg[1]).max() < 1e-9
    rnd = rng.random((1000, P)) * 2 * np.pi
    m2 = engine.magic(qc.run_batch(rnd), (2.0,)).cpu().numpy()[0]
    assert m2.min() > -1e-9 and m2.max() <= np.log((2 ** 12 + 1) / 2) + 1e-9


def test_config3_size_npqc_identity_qfim_16q():
    """The NPQC theorem QFIM(theta_ref) = I (tests.py:103-128) at 16 qubits: an exact
    known answer for the multi-pass derivative pipeline at the headline state size."""
    layers, th = pyqc.templates.NPQC_layers(4, 16)
    qc = pyqc.PQC(16)
    for l in layers:
        qc.add_layer(l)
    F = qc.qfim_batch(np.array([th, th], dtype=np.float64)).cpu().numpy()
    assert np.abs(F - np.eye(F.shape[1])).max() < 1e-12
    qc.update_state(th)
    Q = pyqc.measure.Measurements(qc).get_QFI()
    assert np.abs(Q - np.eye(len(Q))).max() < 1e-12
    assert pyqc.measure.Measurements(qc).get_effective_quantum_dimension(1e-12) == len(th)


# ---- host loops that drive the hot path ---------------------------------
def test_tfim_training_reaches_ground_state():
    """/root/reference/tests.py:130-156: BFGS on a 4-qubit, 4-layer TFIM circuit."""
    import random
    from math import isclose
    N, p = 4, 4
    tfim = pyqc.PQC(N)
    ham = pyqc.templates.TFIM_hamiltonian(N, g=1, h=0)
    e0, ground = ham.groundstate()
    tfim.set_H(ham)
    for l in pyqc.templates.TFIM_layers(p, N):
        tfim.add_layer(l)
    random.seed(11)
    angles = [random.random() * np.pi for _ in range(2 * p)]
    m = pyqc.measure.Measurements(tfim)
    energy, traj, magics, ents, gkps = m.train(method="BFGS", angles=angles)
    assert len(traj) == len(magics) == len(ents) == len(gkps) > 2
    assert isclose(tfim.fidelity(ground), 1, abs_tol=1e-6)
    assert abs(energy - e0) < 1e-6
    # gradient descent branch: energy decreases monotonically for a small rate
    e2, traj2, *_ = m.train(method="gradient", angles=angles, rate=0.01, epsilon=1e-3)
    assert all(b <= a + 1e-12 for a, b in zip(traj2, traj2[1:]))


def test_gradient_vector_matches_finite_differences():
    qc = pyqc.templates.generate_circuit("generic_HE", 5, 2)
    m = pyqc.measure.Measurements(qc)
    th = list(np.random.default_rng(2).random(qc.n_true_params) * 2 * np.pi)
    g = np.array(m.get_gradient_vector(th))
    eps = 1e-6
    for k in (0, 7, 19):
        up, dn = list(th), list(th)
        up[k] += eps
        dn[k] -= eps
        assert abs((qc.cost(up) - qc.cost(dn)) / (2 * eps) - g[k]) < 1e-7


def test_effective_hilbert_space_generic_he():
    """/root/reference/tests.py:297-309 for n = 2, 4 (statistical, rel 0.1 as there)."""
    from math import isclose
    reseed()
    for n in (2, 4):
        qc = pyqc.templates.generate_circuit("generic_HE", n, 2 * n)
        m = pyqc.measure.Measurements(qc)
        f = m._gen_f_samples(150)
        assert isclose(m.find_eff_H(f, n), 2 ** n, rel_tol=0.1)


def test_example_script_flow():
    """/root/reference/example.py end to end (sizes reduced for time)."""
    import random
    ex = cases.build_example4(pyqc)
    cap = pyqc.measure.Measurements(ex)
    reseed()
    assert cap.expressibility(150) > 0 and cap.entropy_of_magic(150) > 0
    random.seed(5)
    ang = [random.random() * np.pi for _ in range(12)]
    out = cap.train(method="BFGS", angles=ang)
    cap.set_minimise_function(cap.theta_to_magic)
    out_magic = cap.train(method="BFGS", angles=ang)
    assert out_magic[2][-1] >= out[2][-1] - 1e-6           # optimising for magic finds more magic
    assert "4 qubit, 3 layer deep PQC" in repr(ex)


def test_find_overparam_point_matches_reference(golden_r2, capsys):
    """measure.py:101-121: layers are appended until the QFIM rank stops growing; the count, the
    printed rank sequence and the circuit it leaves behind equal the reference's (module RNG at
    seed 1, as recorded by tests/golden/make_golden_r2.py)."""
    qc = cases_r2.build_overparam3(pyqc)
    m = pyqc.measure.Measurements(qc)
    reseed()
    count = m.find_overparam_point([0])
    assert count == int(golden_r2["overparam/count"])
    assert qc.n_layers == int(golden_r2["overparam/n_layers_after"])
    assert capsys.readouterr().out == str(golden_r2["overparam/log"])


@pytest.mark.parametrize("method,rate,eps", [("QNG", 0.05, 1e-4), ("gradient", 0.05, 1e-5)])
def test_train_gradient_and_qng_match_reference(golden_r2, method, rate, eps):
    """measure.py:473-553, the fixed-step branches (natural gradient: pinv(QFIM) . grad,
    measure.py:523-529): same number of iterations, same energy / magic / Q / GKP traces."""
    qc = cases_r2.build_tfim3(pyqc)
    m = pyqc.measure.Measurements(qc)
    energy, traj, magics, ents, gkps = m.train(epsilon=eps, rate=rate, method=method,
                                               angles=list(cases_r2.TFIM3_START))
    g = golden_r2
    assert len(traj) == len(g[f"train/{method}/traj"])
    assert abs(energy - float(g[f"train/{method}/energy"])) < 1e-9
    assert np.abs(np.array(traj) - g[f"train/{method}/traj"]).max() < 1e-9
    assert np.abs(np.array(magics) - g[f"train/{method}/magics"]).max() < 1e-8
    assert np.abs(np.array(ents) - g[f"train/{method}/ents"]).max() < 1e-9
    assert np.abs(np.array(gkps) - g[f"train/{method}/gkps"]).max() < 1e-8
    assert np.abs(np.array(qc.get_params()) - g[f"train/{method}/final_angles"]).max() < 1e-9


@pytest.mark.parametrize("n", [4, 6])
def test_effective_hilbert_space_zfsim_matches_reference(golden_r2, n):
    """The zfsim half of tests.py:311-342: fidelity samples of a half-filled, particle-number
    conserving circuit and the Hilbert-space dimension find_eff_H fits to them (about
    C(n, n/2)), against the reference run from the same initial state and RNG position."""
    qc = pyqc.templates.generate_circuit("zfsim", n, n)
    qc.initial_state = golden_r2[f"zfsim/{n}/init"]          # State(array)
    m = pyqc.measure.Measurements(qc)
    reseed()
    F = m._gen_f_samples(60)
    assert np.abs(np.array(F) - golden_r2[f"zfsim/{n}/F"]).max() < ATOL
    eff = m.find_eff_H(F, n)
    ref = float(golden_r2[f"zfsim/{n}/effH"])
    assert abs(eff - ref) < 1e-4 * ref
    from math import comb, isclose
    assert isclose(eff, comb(n, n // 2), rel_tol=0.4)           # the reference's own assertion


@pytest.mark.gpu
def test_nccl_sample_sharding_matches_single_gpu():
    """dist.py over NCCL on 2 GPUs (skipped on a 1-GPU box; the gloo tests cover the logic)."""
    import json
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29517", os.path.join(root, "tools", "dist_nccl_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    assert json.loads(line)["ok"]


def test_reference_side_binding_runs_through_the_c_abi():
    """integration/pqc_b200_binding.py (the stub of INTEGRATION.md B; its `lower` is checked
    against the unmodified reference in tests/test_reference_binding.py) drives libpqc_b200.so
    with nothing but ctypes: same states as the package's own path, Q from pqc_meyer_wallach."""
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    try:
        b = importlib.import_module("pqc_b200_binding")
    finally:
        sys.path.pop(0)
    for kind, n, p in (("generic_HE", 6, 3), ("XXZ", 8, 2), ("fermionic", 4, 1), ("NPQC", 13, 3)):
        qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
        ang = np.random.default_rng(n).random((5, qc.n_true_params)) * 2 * np.pi
        st = b.run_batch(qc, ang)
        assert torch.equal(torch.view_as_real(st), torch.view_as_real(qc.run_batch(ang)))
        assert torch.equal(b.meyer_wallach(st, n), engine.meyer_wallach(st))


def test_arbgate_circuit_matches_reference(golden_r2):
    """ARBGATE = exp(-i theta H) of a dense Hamiltonian (gates.py:407-435) inside a circuit:
    states, cost, every derivative state, QFIM and EQD against the unmodified reference (dense
    expm on the shim); the batched path equals the one-sample path."""
    qc = cases_r2.build_arb4(pyqc)
    assert qc.n_true_params == 13 and "dense gate" in qc.program.describe()
    m = pyqc.measure.Measurements(qc)
    for k, ang in enumerate(cases_r2.ARB4_ANGLES):
        st = qc.run(list(ang))
        assert np.abs(st.numpy() - golden_r2[f"arb4/{k}/state"]).max() < ATOL
        assert abs(qc.cost(list(ang)) - float(golden_r2[f"arb4/{k}/cost"])) < ATOL
        gr = np.stack([g.numpy() for g in qc.get_gradients()])
        assert np.abs(gr - golden_r2[f"arb4/{k}/grads"]).max() < ATOL
        assert rel(m.get_QFI(), golden_r2[f"arb4/{k}/qfi"]) < RTOL
        assert m.get_effective_quantum_dimension(1e-12) == int(golden_r2[f"arb4/{k}/eqd"])
    A = np.array(cases_r2.ARB4_ANGLES)
    stb = qc.run_batch(A).cpu().numpy()
    for k in range(2):
        assert np.abs(stb[k] - golden_r2[f"arb4/{k}/state"]).max() < ATOL
    Fb = qc.qfim_batch(A).cpu().numpy()
    for k in range(2):
        assert rel(Fb[k], golden_r2[f"arb4/{k}/qfi"]) < RTOL
    # the gate's own operator surface
    g = pyqc.ARBGATE(pyqc.templates.TFIM_hamiltonian(3, 0.5))
    g.set_theta(0.37)
    U = g.operation.full()
    assert np.abs(U @ U.conj().T - np.eye(8)).max() < 1e-13


def test_dense_apply_large_register_is_unitary():
    """The dense eigenbasis products at a size where the tiles repeat (10 qubits, 70 states):
    exp(-i theta H) keeps norms, theta = 0 is the identity, and exp(-i a H) exp(-i b H) =
    exp(-i (a + b) H)."""
    n = 10
    H = pyqc.templates.TFIM_hamiltonian(n, 0.9, 0.2)
    qa = pyqc.PQC(n)
    qa.add_layer([pyqc.R_y(i, n) for i in range(n)] + [pyqc.ARBGATE(H), pyqc.ARBGATE(H)])
    qb = pyqc.PQC(n)
    qb.add_layer([pyqc.R_y(i, n) for i in range(n)] + [pyqc.ARBGATE(H)])
    rng = np.random.default_rng(3)
    A = rng.random((70, n + 2)) * 2 * np.pi
    B = np.concatenate([A[:, :n], (A[:, n] + A[:, n + 1])[:, None]], axis=1)
    sa, sb = qa.run_batch(A), qb.run_batch(B)
    assert np.abs((sa - sb).cpu().numpy()).max() < 1e-10
    assert np.abs(engine.overlap(sa, sa).cpu().numpy() - 1).max() < 1e-12
    Z = np.concatenate([A[:, :n], np.zeros((70, 1))], axis=1)
    q0 = pyqc.PQC(n)
    q0.add_layer([pyqc.R_y(i, n) for i in range(n)])
    assert np.abs((qb.run_batch(Z) - q0.run_batch(A[:, :n])).cpu().numpy()).max() < 1e-12


def test_noncommuting_shared_parameter_vs_reference(golden_r3):
    """shared_parameter(commute=False) whose members do not commute (gates.py:458-466: the sum of
    products times the element-wise conjugate of the block): states, cost, every derivative state,
    QFIM and EQD against the unmodified reference (tests/golden/make_golden_r3.py) and the oracle;
    the batched QFIM entry point takes the same literal path."""
    import cases_r3
    from helpers import specs_from_circuit
    for k, ang in enumerate(cases_r3.NONCOMM3_ANGLES):
        qc = cases_r3.build_noncommuting3(pyqc)
        m = pyqc.measure.Measurements(qc)
        st = qc.update_state(list(ang))
        assert np.abs(st.numpy() - golden_r3[f"noncomm3/{k}/state"]).max() < ATOL
        assert abs(qc.cost(list(ang)) - float(golden_r3[f"noncomm3/{k}/cost"])) < ATOL
        gr = np.stack([g.numpy() for g in qc.get_gradients()])
        assert np.abs(gr - golden_r3[f"noncomm3/{k}/grads"]).max() < ATOL
        specs = specs_from_circuit(qc)
        assert np.abs(gr - orc.gradients(specs, 3, np.array([ang]), None)[0]).max() < ATOL
        F = m.get_QFI()
        ref = golden_r3[f"noncomm3/{k}/qfi"]
        assert np.abs(F - ref).max() < RTOL * max(1.0, np.abs(ref).max())
        assert m.get_effective_quantum_dimension(1e-12) == int(golden_r3[f"noncomm3/{k}/eqd"])
        # the derivative of one gate, as the reference API hands it out (circuit.py:149-172)
        g5 = [g for g in qc.gates if g.param_count > 0][5]
        assert np.abs(qc.take_derivative(g5).numpy() - golden_r3[f"noncomm3/{k}/grads"][5]).max() < ATOL
    qc = cases_r3.build_noncommuting3(pyqc)
    Fb, sb = qc.qfim_batch(np.array(cases_r3.NONCOMM3_ANGLES), want_states=True)
    for k in range(2):
        ref = golden_r3[f"noncomm3/{k}/qfi"]
        assert np.abs(Fb[k].cpu().numpy() - ref).max() < RTOL * max(1.0, np.abs(ref).max())
        assert np.abs(sb[k].cpu().numpy() - golden_r3[f"noncomm3/{k}/state"]).max() < ATOL


def test_gate_sums_vs_reference(golden_r3):
    """Gate.__add__ / __radd__ (gates.py:75-85): sums of gate operators, dense and applied."""
    import cases_r3
    N, a, b, c = cases_r3.build_sum_gates(pyqc)
    assert np.abs((a + b).full() - golden_r3["sum/a_plus_b"]).max() < 1e-14
    assert np.abs(((a + b) + c).full() - golden_r3["sum/a_plus_b_plus_c"]).max() < 1e-14
    assert np.abs((a.operation + c).full() - golden_r3["sum/radd"]).max() < 1e-14
    assert np.abs(sum([a, b]).full() - golden_r3["sum/a_plus_b"]).max() < 1e-14
    st = pyqc.State(torch.as_tensor(golden_r3["sum/state_in"], device="cuda"))
    assert np.abs(((a + b) * st).numpy() - golden_r3["sum/applied"]).max() < ATOL
