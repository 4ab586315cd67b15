"""Circuit builders for the fixtures of make_golden_r3.py (the reference package) and the GPU tests
(pyramaterised_b200): shared_parameter blocks whose members do NOT commute, i.e. the general branch
of /root/reference/pyramaterised/gates.py:458-466, and sums of gates (gates.py:75-85).
Public reference API only."""


def build_noncommuting3(pyqc):
    """3 qubits, 7 parameters: R_x and R_z of the same qubit under one angle, two R_y under one
    angle (non-symmetric matrices: the block's element-wise conjugate is not its inverse), an
    R_zz + R_x pair, between ordinary rotations and entanglers."""
    N = 3
    c = pyqc.PQC(N)
    c.add_layer([pyqc.R_y(i, N) for i in range(N)] + [pyqc.CHAIN(pyqc.CNOT, N)])
    c.add_layer([pyqc.shared_parameter([pyqc.R_x(0, N), pyqc.R_z(0, N)], N, commute=False),
                 pyqc.shared_parameter([pyqc.R_y(0, N), pyqc.R_y(1, N)], N, commute=False),
                 pyqc.CHAIN(pyqc.CPHASE, N)])
    c.add_layer([pyqc.shared_parameter([pyqc.R_zz((0, 1), N), pyqc.R_x(1, N), pyqc.R_y(2, N)], N,
                                       commute=False),
                 pyqc.R_x(2, N)])
    return c


NONCOMM3_ANGLES = [[0.3, 1.1, 2.5, 4.0, 0.21, 5.2, 0.9],
                   [5.9, 0.2, 3.1, 1.4, 1.05, 2.8, 4.4]]


def build_sum_gates(pyqc):
    """Gates whose sums are formed (gates.py:75-85)."""
    N = 3
    a = pyqc.R_x(0, N)
    a.set_theta(0.7)
    b = pyqc.R_zz((0, 2), N)
    b.set_theta(1.9)
    c = pyqc.CNOT([1, 2], N)
    return N, a, b, c
