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
