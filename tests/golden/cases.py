"""Circuit builders shared by the golden generator (run against the *reference*
package) and the parity tests (run against ``pyramaterised_b200``).

Every builder takes the package module as ``pyqc`` and uses only the public
reference API (circuit.py:16-72, templates.py:361-429), so the very same calls
build the circuit in both implementations.
"""
import numpy as np

QG_ANGLES = [3.21587011, 5.97193953, 0.90578156, 5.96054027, 1.9592948, 2.65983852,
             5.20060878, 2.571074, 3.45319898, 0.17315902, 4.73446249, 3.38125416]
QG_ENERGY = 0.46135870050914374          # /root/reference/tests.py:207-209
QG_EQD = 12                              # /root/reference/tests.py:211-212


def build_qg4(pyqc):
    """The 4-qubit quantum-geometry circuit of /root/reference/tests.py:161-190."""
    c = pyqc.PQC(4)
    c.add_layer([pyqc.fixed_R_y(i, 4, np.pi / 4) for i in range(4)])
    c.add_layer([pyqc.R_z(0, 4), pyqc.R_x(1, 4), pyqc.R_y(2, 4), pyqc.R_z(3, 4),
                 pyqc.CHAIN(pyqc.CNOT, 4)])
    c.add_layer([pyqc.R_x(0, 4), pyqc.R_x(1, 4), pyqc.R_x(2, 4), pyqc.R_y(3, 4),
                 pyqc.CHAIN(pyqc.CNOT, 4)])
    c.add_layer([pyqc.R_z(0, 4), pyqc.R_x(1, 4), pyqc.R_y(2, 4), pyqc.R_y(3, 4),
                 pyqc.CHAIN(pyqc.CNOT, 4)], n=1)
    return c


def build_example4(pyqc):
    """/root/reference/example.py:7-13 -- R_x layer + CNOT chain, three times."""
    N = 4
    c = pyqc.PQC(N)
    c.add_layer([pyqc.R_x(i, N) for i in range(N)] + [pyqc.CHAIN(pyqc.CNOT, N)], n=3)
    return c


def build_mixed5(pyqc):
    """Hand-made circuit touching the fixed gates, CZ/CPHASE, ALLTOALL and every RR."""
    N = 5
    c = pyqc.PQC(N)
    c.add_layer([pyqc.H(0, N), pyqc.X(1, N), pyqc.S(2, N), pyqc.T(3, N),
                 pyqc.fixed_R_z(4, N, 0.3), pyqc.fixed_R_y(2, N, 1.1)])
    c.add_layer([pyqc.R_y(i, N) for i in range(N)] + [pyqc.CZ([0, 3], N), pyqc.CPHASE([4, 1], N),
                                                      pyqc.CNOT([3, 1], N), pyqc.CNOT([0, 4], N)])
    c.add_layer([pyqc.R_xx([0, 2], N), pyqc.R_yy([3, 1], N), pyqc.R_zz([4, 2], N),
                 pyqc.R_x(2, N), pyqc.ALLTOALL(pyqc.CZ, N)])
    c.add_layer([pyqc.R_z(i, N) for i in range(N)] + [pyqc.CHAIN(pyqc.CPHASE, N),
                                                      pyqc.RR_block(pyqc.R_yy, N)])
    return c


def build_template(kind, N, p, **kw):
    def _b(pyqc):
        return pyqc.templates.generate_circuit(kind, N, p, shuffle=False, **kw)
    return _b


# name -> (builder, n_samples, n_grad_samples, want_magic)
CASES = {
    "qg4":            (build_qg4, 4, 2, True),
    "example4":       (build_example4, 4, 1, True),
    "mixed5":         (build_mixed5, 4, 2, True),
    "npqc_4_4":       (build_template("NPQC", 4, 4), 8, 2, True),
    "npqc_6_5":       (build_template("NPQC", 6, 5), 3, 1, True),
    "he_5_3":         (build_template("generic_HE", 5, 3), 24, 1, True),
    "he_10_10":       (build_template("generic_HE", 10, 10), 3, 0, False),
    "clifford_4_2":   (build_template("clifford", 4, 2), 3, 1, True),
    "tfim_4_4":       (build_template("TFIM", 4, 4), 4, 2, True),
    "tfim_6_3":       (build_template("TFIM", 6, 3), 3, 2, True),
    "tfimmod_4_2":    (build_template("TFIM_modified", 4, 2), 3, 1, True),
    "xxz_4_2":        (build_template("XXZ", 4, 2), 4, 2, True),
    "xxz_6_2":        (build_template("XXZ", 6, 2), 3, 2, True),
    "circuit1_4_2":   (build_template("Circuit_1", 4, 2), 3, 1, False),
    "circuit2_5_2":   (build_template("Circuit_2", 5, 2), 3, 1, False),
    "circuit9_4_2":   (build_template("Circuit_9", 4, 2), 3, 1, False),
    "ycphase_5_2":    (build_template("y_CPHASE", 5, 2), 3, 1, False),
    "dycphase_4_2":   (build_template("double_y_CPHASE", 4, 2), 3, 1, False),
    "qgt_4_2":        (build_template("qg_circuit", 4, 2), 3, 1, False),
    # SURVEY.md 8(f) "next" rows -- recorded now so the later widening has its fixtures
    "fermionic_4_1":  (build_template("fermionic", 4, 1), 3, 1, False),
    "fermionic_6_1":  (build_template("fermionic", 6, 1), 2, 1, False),
    "zfsim_4_2":      (build_template("zfsim", 4, 2), 3, 1, False),
    "fsim_5_2":       (build_template("fsim", 5, 2, rotator="x"), 3, 1, False),
    "fixedfsim_4_2":  (build_template("fixed_fsim", 4, 2), 3, 1, False),
}


def n_true_params(circuit):
    """True parameter count (circuit.n_params is 2x this, quirk Q1, circuit.py:62-72)."""
    return len(circuit.get_params())


def case_angles(name, circuit, S):
    seed = 1000 + sorted(CASES).index(name)
    P = n_true_params(circuit)
    return np.random.default_rng(seed).random((S, P)) * 2 * np.pi
