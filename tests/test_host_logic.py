"""CPU-only checks of the host side: reference-API bookkeeping, templates, lowering, the
C-ABI surface, and that the product fails loudly instead of falling back to a CPU path."""
import os
import re

import numpy as np
import pytest

import cases
import pyramaterised_b200 as pyqc
from helpers import oracle_case, specs_from_circuit
from oracle import pqc_oracle as orc
from pyramaterised_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL = sorted(cases.CASES)


@pytest.mark.parametrize("name", ALL)
def test_templates_match_oracle_specs(golden, name):
    """The same builder calls as the golden generator produce the gate list the oracle
    restates from templates.py, and the quirk-Q1 bookkeeping of circuit.py:62-72."""
    qc = cases.CASES[name][0](pyqc)
    specs, n, init = oracle_case(name)
    assert qc.n_qubits == n
    assert specs_from_circuit(qc) == specs
    assert qc.n_true_params == int(golden[f"{name}/P"]) == cases.n_true_params(qc)
    assert qc.n_params == int(golden[f"{name}/n_params_attr"])
    assert qc.parameterised == list(golden[f"{name}/parameterised"])
    if init is not None:
        assert int(np.argmax(np.abs(init))) == qc._init_index


def test_lowering_slots_and_groups():
    qc = pyqc.templates.generate_circuit("TFIM", 16, 16)
    ops = qc.lower()
    assert len(ops) == 16 + 16 * 32 and qc.n_true_params == 32
    slots = [o[3] for o in ops if o[3] >= 0]
    assert slots == sorted(slots) and set(slots) == set(range(32))
    assert [o[5] for o in ops] == sorted(o[5] for o in ops)          # gate index is monotone
    fer = pyqc.templates.generate_circuit("fermionic", 4, 1, shuffle=False).lower()
    neg = [o for o in fer if o[0] == _lib.OP_RZ and o[6] == -1.0]
    off = [o for o in fer if o[0] == _lib.OP_RZ and o[3] >= 0 and o[7] == np.pi]
    assert neg and off and neg[0][3] == off[0][3]                     # shared slot
    fs = pyqc.templates.generate_circuit("fsim", 5, 2, shuffle=False).lower()
    two = [o for o in fs if o[0] == _lib.OP_FSIM]
    assert all(o[4] == o[3] + 1 for o in two)


def test_random_stream_matches_reference_rng():
    """gates.py:10 + circuit.py:100-112: one rng.random(1) per parameter, sample-major."""
    qc = pyqc.templates.generate_circuit("NPQC", 4, 4)
    st = pyqc.gates.rng.bit_generator.state
    try:
        pyqc.gates.rng.bit_generator.state = np.random.default_rng(1).bit_generator.state
        batch = qc.draw_random(3)
        pyqc.gates.rng.bit_generator.state = np.random.default_rng(1).bit_generator.state
        single = [qc.set_params("random").copy() for _ in range(3)]
    finally:
        pyqc.gates.rng.bit_generator.state = st
    ref = np.random.default_rng(1).random(3 * 20) * 2 * np.pi
    assert np.array_equal(batch.ravel(), ref) and np.array_equal(np.concatenate(single), ref)


def test_set_params_errors_like_reference():
    qc = pyqc.templates.generate_circuit("generic_HE", 3, 1)
    with pytest.raises(Exception, match="No parameters supplied!"):
        qc.set_params("nonsense")
    with pytest.raises(IndexError):
        qc.set_params([0.1, 0.2])
    with pytest.raises(Exception, match="Must supply a valid entangler!"):
        pyqc.templates.string_to_entangler("swap")
    qc.set_params(list(range(6)))
    assert qc.get_params() == list(range(6))


def test_pauli_sum_algebra():
    qt = pyqc.qt
    X, Y, Z = qt.sigmax(), qt.sigmay(), qt.sigmaz()
    assert X * Y == 1j * Z and Y * Z == 1j * X and Z * X == 1j * Y
    H = pyqc.templates.TFIM_hamiltonian(4, g=1.0, h=0.5)
    M = H.full()
    assert np.allclose(M, M.conj().T) and H.isherm
    ref = np.zeros((16, 16), complex)
    P = {"x": orc.SX, "z": orc.SZ}

    def emb(op, q):
        m = np.array([[1.0]])
        for k in range(4):
            m = np.kron(m, op if k == q else np.eye(2))
        return m
    for i in range(4):
        ref -= emb(P["z"], i) @ emb(P["z"], (i + 1) % 4) + emb(P["x"], i) + 0.5 * emb(P["z"], i)
    assert np.allclose(M, ref)
    zz = pyqc.PQC(3).H
    assert np.allclose(zz.full(), emb(orc.SZ, 0)[:8, :8] * 0 + np.kron(np.kron(orc.SZ, orc.SZ), np.eye(2)))
    assert (Y.conj() == -1 * Y) and (qt.tensor([X, Z]).n == 2)


def test_shared_parameter_derivative_shortcut_conditions():
    N = 4
    xxz = pyqc.templates.XXZ_layers(1, N)[0]
    assert all(g._sum_of_generators_is_exact() for g in xxz)
    bad = pyqc.shared_parameter([pyqc.R_x(0, N), pyqc.R_z(0, N)], N, commute=False)
    assert not bad._sum_of_generators_is_exact()
    bad2 = pyqc.shared_parameter([pyqc.R_y(0, N), pyqc.R_y(1, N)], N, commute=False)
    assert not bad2._sum_of_generators_is_exact()
    d = pyqc.RR_block(pyqc.R_zz, N).derivative()
    assert len(d.terms) == N and all(abs(c + 0.5j) < 1e-15 for c in d.terms.values())


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pqc_b200.h")).read()
    declared = set(re.findall(r"PQC_API\s+(?:const\s+char\*|int|long long)\s+(pqc_\w+)\s*\(", hdr))
    assert len(declared) >= 23
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared - {"pqc_last_error", "pqc_launch_count"} == set(_lib.SIGNATURES)
    assert lib.pqc_abi_version() == 1
    # error path without touching the device
    import ctypes as C
    h = C.c_void_p()
    assert lib.pqc_program_create(0, 0, 0, None, C.byref(h)) < 0
    assert b"n_qubits" in lib.pqc_last_error()


def test_c_abi_from_a_plain_c_host(tmp_path):
    """The boundary is a C ABI, not a Python one: include/pqc_b200.h compiles as pedantic C99 and a
    C host linked against libpqc_b200.so plans a gate program (TFIM-like layer on 16 qubits: H, R_zz
    ring, R_x -- gates.py:106-147,493-537) without a GPU, reads the plan back and gets the library's
    error string on a bad call.  No compute entry point is called."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "host.c"
    src.write_text(r"""
#include "pqc_b200.h"
#include <stdio.h>
#include <string.h>
int main(void) {
  enum { N = 16 };
  pqc_op ops[3 * N];
  int k = 0, q;
  pqc_program* prog = 0;
  int64_t st[8];
  static char text[1 << 16];
  for (q = 0; q < N; ++q, ++k) {            /* Hadamard layer */
    pqc_op o = {PQC_OP_H, 0, -1, -1, -1, 0, 1.0, 0.0};
    o.q0 = q; o.group = k; ops[k] = o;
  }
  for (q = 0; q < N; ++q, ++k) {            /* R_zz ring sharing parameter 0 */
    pqc_op o = {PQC_OP_RZZ, 0, 0, 0, -1, 0, 1.0, 0.0};
    o.q0 = q; o.q1 = (q + 1) % N; o.group = N; ops[k] = o;
  }
  for (q = 0; q < N; ++q, ++k) {            /* R_x layer sharing parameter 1 */
    pqc_op o = {PQC_OP_RX, 0, -1, 1, -1, 0, 1.0, 0.0};
    o.q0 = q; o.group = N + 1; ops[k] = o;
  }
  if (pqc_abi_version() != 1) return 2;
  if (pqc_program_create(N, 2, k, ops, &prog) != 0) { printf("create: %s\n", pqc_last_error()); return 3; }
  if (pqc_program_stats(prog, st) != 0) return 4;
  if (pqc_program_describe(prog, text, sizeof text) < 0) return 5;
  printf("n=%lld P=%lld ops=%lld passes=%lld grad=%lld\n", (long long)st[0], (long long)st[1],
         (long long)st[2], (long long)st[3], (long long)st[5]);
  printf("has_pass=%d\n", strstr(text, "PASS") != 0);
  pqc_program_destroy(prog);
  prog = 0;
  if (pqc_program_create(0, 0, 0, 0, &prog) >= 0) return 6;
  printf("err=%s\n", pqc_last_error());
  return 0;
}
""")
    exe = tmp_path / "host"
    libdir = os.path.join(ROOT, "pyramaterised_b200")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror",
                    "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lpqc_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert lines[0].startswith("n=16 P=2 ops=48 passes=") and lines[0].endswith("grad=1")
    assert 1 <= int(re.search(r"passes=(\d+)", lines[0]).group(1)) <= 3
    assert lines[1] == "has_pass=1" and "n_qubits" in lines[2]


def test_no_cpu_fallback_and_no_oracle_import():
    import torch
    pkg = os.path.join(ROOT, "pyramaterised_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src and "qutip_lite" not in src, fn
    if not torch.cuda.is_available():
        qc = pyqc.templates.generate_circuit("generic_HE", 3, 1)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            qc.run("random")


def test_planner_runs_without_gpu_and_fuses_tfim_layers():
    """Planning is host-only.  TFIM-16x16 must come out as ONE pass per layer (17 with the
    Hadamard layer) + one gather per R_x parameter, every pass loading and storing directly."""
    qc = pyqc.templates.generate_circuit("TFIM", 16, 16)
    prog = qc.program
    lines = prog.describe().splitlines()
    passes = [l for l in lines if l.startswith("PASS")]
    gathers = [l for l in lines if l.startswith("GATHER")]
    assert prog.n_qfim_passes == 17 and len(gathers) == 16
    assert all("direct=1/1" in l for l in passes)
    assert all("fast=1" in l for l in passes)                  # aligned nibble sweeps: k_layer_pass
    assert sum("spawns=1" in l for l in passes) == 16          # R_zz rings: in-pass diagonal spawns
    assert all(l.count("[rb") <= 3 for l in passes[:17])        # three sweeps per pass
    # the 256 R_zz of the circuit became 16 fused phase ops (internal opcode 32)
    assert sum(l.count(" 32") for l in passes) == 16
    xxz = pyqc.templates.generate_circuit("XXZ", 16, 16, shuffle=False).program
    assert xxz.n_qfim_passes <= 48 and " 33" in xxz.describe()   # XY fusion (opcode 33)
    small = pyqc.templates.generate_circuit("generic_HE", 10, 10).program
    assert small.n_passes <= 22 and small.grad_supported
    tiny = pyqc.templates.generate_circuit("NPQC", 4, 4).program     # n < 8: v0 tile plan
    assert "v0 plan" in tiny.describe()


def test_planner_meet_in_the_middle_qfim_plan():
    """QFIM plan: the circuit is cut in the middle, the second half is differentiated backwards
    from psi(T); the vector-pass count (what the HBM traffic is proportional to) drops ~1.6x."""
    import re
    d = pyqc.templates.generate_circuit("TFIM", 16, 16).program.describe()
    m = re.search(r"BIDIR cut=(\d+)/(\d+) PF=(\d+) PB=(\d+) vector-passes (\d+) -> (\d+)", d)
    assert m, d[-300:]
    cut, nops, pf, pb, fwd, bi = map(int, m.groups())
    assert (cut, nops, pf, pb) == (272, 528, 16, 16)          # H layer + 8 of 16 layers
    assert fwd == 289 and bi <= 185
    # circuits with a gate that is not inverted by negating its angle keep the forward plan
    assert "BIDIR" not in pyqc.templates.generate_circuit("fsim", 12, 2).program.describe()
    # too few parameters / qubits: forward plan
    assert "BIDIR" not in pyqc.templates.generate_circuit("TFIM", 8, 4).program.describe()


def test_planner_layer_sequence_plans():
    """XXZ (XY pair rotations, nibble sweeps in any order, staged ends) and NPQC (runs of R_z /
    CZ, many layers per pass) passes are lowered to SeqPlans for k_layer_seq (describe: fast=2);
    TFIM keeps the layer pass (fast=1); CNOT-chain circuits stay on the generic kernel."""
    import re

    def tally(kind, n, p, prefix):
        d = pyqc.templates.generate_circuit(kind, n, p, shuffle=False).program.describe()
        lines = [l for l in d.splitlines() if l.startswith(prefix)]
        return {k: sum(("fast=%d" % k) in l for l in lines) for k in (0, 1, 2)}, lines

    xxz, lines = tally("XXZ", 16, 16, "PASS")
    assert xxz[0] == 0 and xxz[1] == 0 and xxz[2] == len(lines) > 40
    assert any("direct=0/0" in l for l in lines)               # staged ends are covered
    run, _ = tally("XXZ", 16, 16, "RUN PASS")
    assert run[0] == 0 and run[2] >= 40
    npqc, lines = tally("NPQC", 28, 20, "RUN PASS")
    assert npqc[0] == 0 and npqc[1] + npqc[2] == len(lines) == 41
    packed, lines = tally("NPQC", 16, 16, "RUN PASS")           # 16 layers in 5 passes
    assert packed[0] == 0 and len(lines) <= 6
    assert max(int(re.search(r"mops=(\d+)", l).group(1)) for l in lines) > 72
    tfim, lines = tally("TFIM", 16, 16, "PASS")
    assert tfim[1] == len(lines) and tfim[2] == 0
    he, _ = tally("generic_HE", 14, 2, "PASS")
    assert he[2] == 0                                           # CNOT permutations: generic kernel


@pytest.mark.parametrize("n", [1, 3, 5])
def test_conversion_matrices_match_reference(golden_r2, n):
    """measure.py:268-316: the xor / sign lookup tables of the reference's dense magic
    formulation (host code; kept for API compatibility) against reference-generated fixtures."""
    m = pyqc.measure.Measurements(pyqc.PQC(n))
    xor, sign = m.get_conversion_matrices()
    assert np.array_equal(np.asarray(xor), golden_r2[f"conv/{n}/xor"])
    assert np.array_equal(np.asarray(sign), golden_r2[f"conv/{n}/sign"])
    assert np.array_equal(m.numberToBase(3 % (2 ** n), 2, n), golden_r2[f"conv/{n}/base3"])
    ox, osg = orc.conversion_matrices(n)
    assert np.array_equal(np.asarray(xor), ox) and np.array_equal(np.asarray(sign), osg)
    m.set_converstion_matrices((xor, sign))
    assert m.conversion_matrices[0] is xor


def test_arbgate_host_surface():
    """ARBGATE bookkeeping without a GPU (gates.py:407-435): one parameter, Hermitian check,
    operation = exp(-i theta H) (NOT theta / 2), derivative = -i H / 2, circuits that contain it
    plan as segments."""
    import scipy.linalg
    import pyramaterised_b200 as pyqc
    H = pyqc.templates.TFIM_hamiltonian(3, 0.8, 0.1)
    g = pyqc.ARBGATE(H)
    assert g.is_param and g.param_count == 1 and g.q_N == 3
    g.set_theta(0.6)
    assert np.abs(g.operation.full() - scipy.linalg.expm(-0.6j * H.full())).max() < 1e-13
    assert np.abs(g.derivative().full() - (-0.5j * H.full())).max() < 1e-15
    with pytest.raises(ValueError):
        pyqc.ARBGATE(np.array([[0, 1], [0, 0]], dtype=complex))
    qc = pyqc.PQC(3)
    qc.add_layer([pyqc.R_x(0, 3), pyqc.ARBGATE(H), pyqc.R_z(2, 3)])
    assert qc.n_true_params == 3 and qc.n_params == 6          # quirk Q1 bookkeeping unchanged
    prog = qc.program
    assert [s[0] for s in prog.segs] == ["ops", "dense", "ops"] and prog.P == 3


def test_fsim_matrix_functions_and_get_op_surface():
    """The module-level fSim matrices of gates.py:588-648, 700-737 (quirk-Q3 entries included), their
    embedding on a register (qutip's gate_expand_2toN, qubit 0 most significant, first factor =
    control) against the oracle's matrix-free two-qubit application, and the get_op / set_properties
    surface every reference gate class has (gates.py:122-127)."""
    G = pyqc.gates
    th, ph = 0.37, 1.21
    for d, fn in enumerate((G.fsim_gate, G.fsim_gate_d_theta, G.fsim_gate_d_phi)):
        assert np.array_equal(fn(th, ph).full(), orc.fsim_matrix(th, ph, d))
    assert np.array_equal(G.fixed_fsim_gate(th).full(), orc.fixed_fsim_matrix(th, 0))
    assert np.array_equal(G.fixed_fsim_gate_d_theta(th).full(), orc.fixed_fsim_matrix(th, 1))
    assert G.fsim_gate_d_phi(th, ph).dims == [[2, 2], [2, 2]]
    # control = 1, target = 0 without N means N = 2 (gates.py:591-592)
    assert G.fsim_gate_d_theta(th, ph, control=1, target=0).shape == (4, 4)
    rng = np.random.default_rng(5)
    for (N, c, t) in [(2, 1, 0), (3, 0, 2), (4, 3, 1), (5, 1, 2)]:
        psi = rng.standard_normal((1, 1 << N)) + 1j * rng.standard_normal((1, 1 << N))
        for d, fn in ((1, G.fsim_gate_d_theta), (2, G.fsim_gate_d_phi)):
            M = fn(th, ph, N=N, control=c, target=t)
            assert M.dims == [[2] * N, [2] * N]
            want = orc.apply_2q(psi.copy(), N, c, t, orc.fsim_matrix(th, ph, d))[0]
            assert np.abs(M.full() @ psi[0] - want).max() < 1e-14
        M = G.fixed_fsim_gate_d_theta(th, N=N, control=c, target=t)
        want = orc.apply_2q(psi.copy(), N, c, t, orc.fixed_fsim_matrix(th, 1))[0]
        assert np.abs(M.full() @ psi[0] - want).max() < 1e-14
    # on a register the gates themselves are one-op symbolic operators (any N)
    op = G.fsim_gate(th, ph, N=20, control=3, target=4)
    assert op.n == 20 and [o[0] for o in op.ops] == [_lib.OP_FSIM]
    assert [o[0] for o in G.fixed_fsim_gate(th, N=20, control=3, target=4).ops] == [_lib.OP_FIXED_FSIM]
    with pytest.raises(ValueError):
        G.fsim_gate_d_phi(th, ph, N=3, control=1, target=1)
    # gate objects: derivative operators as the reference returns them
    f = pyqc.fSim([0, 2], 3)
    f.set_theta(th)
    f.set_phi(ph)
    assert f.parameterised_derivative(1) == G.fsim_gate_d_theta(th, ph, N=3, control=0, target=2)
    assert f.parameterised_derivative(2) == G.fsim_gate_d_phi(th, ph, N=3, control=0, target=2)
    with pytest.raises(AttributeError):
        f.derivative()
    ff = pyqc.fixed_fSim([1, 0], 2)
    ff.set_theta(th)
    assert ff.derivative() == G.fixed_fsim_gate_d_theta(th, N=2, control=1, target=0)
    for g in (pyqc.R_x(0, 2), pyqc.H(1, 2), pyqc.CNOT([0, 1], 2), pyqc.CHAIN(pyqc.CZ, 3),
              pyqc.R_zz([0, 1], 2), f, ff):
        assert g.set_properties() is None
        assert g.get_op().ops == g.operation.ops


def test_front_planner_plans_and_plan_choice():
    """The light-cone planner (pqc_front.cu), host only.  NPQC: the odd qubits (one first-layer
    rotation each, then only CZ partners: cheap unblockers) are cleared in light sweeps first, so
    the even qubits' chains run with all four register slots busy and every CZ merged into its
    rotation -- 16q x 16 layers in 2 passes and ~4 layer ops per layer instead of 11.  Plan choice
    for PQC.run: front plan for CNOT-chain and NPQC circuits, block plan for TFIM / XXZ."""
    import re
    import pyramaterised_b200 as pyqc

    def front(kind, n, p):
        qc = pyqc.templates.generate_circuit(kind, n, p, shuffle=False)
        lines = qc.program.describe().split("\n")
        head = [l for l in lines if l.startswith("FRONT plan")]
        ops = []
        for l in lines:
            if l.startswith("FRONT PASS"):
                for _, o in re.findall(r"\[rb ([0-9,]+) pre\d+ post\d+:([0-9 ]*)\]", l):
                    ops += o.split()
        passes = sum(l.startswith("FRONT PASS") for l in lines)
        return head[0], passes, ops, qc

    head, passes, ops, qc = front("NPQC", 16, 16)
    assert head.endswith("used for PQC.run: 1") and passes == 2 and qc.program.n_passes == 2
    layer_ops = sum(o in ("35", "36", "37", "38") for o in ops)
    assert layer_ops <= 80, layer_ops                     # 185 before the unblocker sweeps
    assert "8" not in ops                                 # no CZ left between two register bits
    head, passes, ops, qc = front("NPQC", 28, 20)
    assert passes <= 6 and sum(o in ("37", "38") for o in ops) <= 200
    head, passes, _, qc = front("generic_HE", 16, 16)
    assert head.endswith("used for PQC.run: 1") and qc.program.n_passes == passes
    head, passes, _, qc = front("TFIM", 16, 16)            # every block pass is a layer pass
    assert head.endswith("used for PQC.run: 0") and qc.program.n_passes > passes
    # XXZ: the front plan (register-bond R_zz as direct phases, op kinds 41 / 42; measured faster than
    # the block plan's light layer-sequence passes from 12 to 28 qubits, profiles/r2_relabel_tables.md)
    head, passes, ops, qc = front("XXZ", 16, 16)
    assert head.endswith("used for PQC.run: 1") and qc.program.n_passes == passes == 10
    assert sum(o in ("41", "42") for o in ops) > sum(o == "32" for o in ops)
    head, passes, _, qc = front("XXZ", 24, 2)
    assert head.endswith("used for PQC.run: 1") and qc.program.n_passes == passes


def test_front_planner_worklist_equals_scan_and_plans_fast(monkeypatch):
    """The front planner simulates candidate tiles and sweeps thousands of times per pass.  Both
    simulations run from the ready frontier (a worklist / a min-heap over ready ops) instead of
    scanning the whole pending list; PQC_FRONT_CHECK=1 runs the original scans next to them at every
    call and aborts the process on any difference.  Planning a deep 12-qubit CNOT-chain circuit
    (1392 ops; 17 s with the scans, meet-in-the-middle sub-plans included) must stay interactive."""
    import time
    t0 = time.time()
    assert "FRONT plan" in pyqc.templates.generate_circuit("qg_circuit", 12, 20).program.describe()
    assert time.time() - t0 < 8.0
    d6 = pyqc.templates.generate_circuit("qg_circuit", 12, 6).program.describe()
    monkeypatch.setenv("PQC_FRONT_CHECK", "1")
    assert pyqc.templates.generate_circuit("qg_circuit", 12, 6).program.describe() == d6
    for kind, n, p in (("NPQC", 16, 16), ("XXZ", 16, 16), ("generic_HE", 16, 16), ("NPQC", 28, 20),
                       ("Circuit_9", 14, 5), ("TFIM", 18, 4), ("clifford", 13, 8)):
        kw = {"shuffle": False} if kind == "XXZ" else {}
        assert "FRONT plan" in pyqc.templates.generate_circuit(kind, n, p, **kw).program.describe()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver runs beside ours): one JSON line with the
    contract's keys, at a size a CPU finishes instantly; it must not need a GPU or the CUDA library."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--qubits", "6",
                        "--layers", "2", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
