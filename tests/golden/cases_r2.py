"""Circuit builders for the round-2 fixtures (make_golden_r2.py runs them against the
reference package, the GPU tests against pyramaterised_b200); public reference API only."""
import numpy as np

TFIM3_START = [0.7, 1.9, 2.6, 0.4, 3.3, 1.2]


def build_overparam3(pyqc):
    """3 qubits, one layer of R_y + R_z rotations and a CNOT chain; find_overparam_point([0])
    keeps appending copies of that layer until the QFIM rank stops growing
    (/root/reference/pyramaterised/measure.py:101-121)."""
    N = 3
    c = pyqc.PQC(N)
    c.add_layer([pyqc.R_y(i, N) for i in range(N)] + [pyqc.R_z(i, N) for i in range(N)] +
                [pyqc.CHAIN(pyqc.CNOT, N)])
    return c


def build_tfim3(pyqc):
    """3-qubit, 3-layer TFIM ansatz with the TFIM Hamiltonian (g = 1) as the cost
    (/root/reference/tests.py:130-156 at a size a fixed-step optimiser finishes quickly)."""
    N, p = 3, 3
    c = pyqc.PQC(N)
    for l in pyqc.templates.TFIM_layers(p, N):
        c.add_layer(l)
    c.set_H(pyqc.templates.TFIM_hamiltonian(N, 1))
    return c


def build_arb4(pyqc):
    """4 qubits with two ARBGATEs (exp(-i theta H), /root/reference/pyramaterised/gates.py:407-435)
    between ordinary layers: H1 = a TFIM Hamiltonian with a longitudinal field, H2 = another
    TFIM plus a Y X string.  13 parameters."""
    N = 4
    qt = pyqc.gates.qt
    f = pyqc.gates.genFockOp
    H1 = pyqc.templates.TFIM_hamiltonian(N, 0.7, 0.3)
    H2 = pyqc.templates.TFIM_hamiltonian(N, 1.3) + 0.5 * f(qt.sigmay(), 1, N) * f(qt.sigmax(), 3, N)
    c = pyqc.PQC(N)
    c.add_layer([pyqc.R_y(i, N) for i in range(N)] + [pyqc.ARBGATE(H1), pyqc.CHAIN(pyqc.CNOT, N)])
    c.add_layer([pyqc.R_x(i, N) for i in range(N)] + [pyqc.ARBGATE(H2), pyqc.R_z(1, N),
                                                      pyqc.fixed_R_y(2, N, 0.4)])
    c.add_layer([pyqc.R_zz([0, 2], N), pyqc.ARBGATE(H1)])
    return c


ARB4_ANGLES = [[0.3, 1.1, 2.5, 4.0, 0.21, 5.2, 0.9, 3.3, 1.7, 0.45, 2.2, 6.0, 0.13],
               [5.9, 0.2, 3.1, 1.4, 1.05, 2.8, 4.4, 0.6, 5.0, 0.8, 3.9, 1.2, 0.66]]
