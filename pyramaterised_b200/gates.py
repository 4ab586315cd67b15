"""Gate classes with the reference's names, constructor signatures and attributes
(/root/reference/pyramaterised/gates.py), re-designed for a matrix-free GPU engine.

A reference gate owns a 2^n x 2^n ``qt.Qobj`` in ``.operation`` that is rebuilt on every
``set_theta`` (gates.py:122-131).  Here a gate is a light description that *lowers* to
primitive ops of the CUDA library (include/pqc_b200.h ``pqc_op``); ``.operation`` is a
symbolic ``Operator`` built on demand, and ``derivative()`` is a symbolic ``PauliSum``.
"""
import operator
from copy import copy, deepcopy                      # noqa: F401  (re-exported like the reference)
from functools import reduce
from itertools import permutations
from typing import Literal, Tuple, Type, Union       # noqa: F401

import numpy as np

from . import _lib
from . import qobj as qt                              # the reference exposes `qt` (gates.py:1)
from .qobj import Composite, Operator, OpSum, PauliSum, State

rng = np.random.default_rng(1)                        # gates.py:10 -- same global stream

QuantumGate = Union["Gate", Operator, PauliSum, State]
DoubleParamGate = "fSim"
Gradient = State
QubitIndex = int
QubitList = Union[list, tuple]
QubitNumber = int
Angle = Union[int, float]
Layer = list
RotationLayer = list
EntanglingLayer = list


def prod(factors):
    """gates.py:30-31"""
    return reduce(operator.mul, factors, 1)


def flatten(l):
    """gates.py:34-35"""
    return [item for sublist in l for item in sublist]


def genFockOp(op, position, size, levels=2, opdim=0):
    """Embed a one-qubit operator at `position` of `size` qubits (gates.py:39-42)."""
    ops = [qt.qeye(levels) for _ in range(size - opdim)]
    ops[position] = op
    return qt.tensor(ops)


def iden(N):
    """gates.py:45-46"""
    return PauliSum.identity(N)


def _op(kind, q0, q1=-1, param=-1, param2=-1, scale=1.0, offset=0.0):
    return (kind, q0, q1, param, param2, 0, float(scale), float(offset))


class Gate:
    """Base class: multiplication / addition go through the symbolic operator, as the
    reference routes them through ``.operation`` (gates.py:49-100)."""

    param_count = 0
    is_param = False

    def __init__(self, q_N):
        self.q_N = q_N
        self.theta = 0
        self.phi = 0

    # ---- lowering ---------------------------------------------------------------------
    def _lower(self, slot):
        """Primitive ops of this gate; `slot` is its first parameter slot (or -1 to
        freeze the current angles)."""
        return []

    def _frozen(self):
        """Primitive ops with the gate's current angles baked in (no parameter slots)."""
        return []

    @property
    def operation(self):
        return Operator(self.q_N, self._frozen())

    # ---- reference protocol (gates.py:63-100) -------------------------------------------
    def __mul__(self, b):
        return self.operation * (b.operation if isinstance(b, Gate) else b)

    def __rmul__(self, b):
        return (b.operation if isinstance(b, Gate) else b) * self.operation

    def __add__(self, b):
        # gates.py:75-79: `self.operation + b.operation` -- a sum of unitaries stays a term list
        return OpSum([self.operation, b.operation if isinstance(b, Gate) else b])

    def __radd__(self, b):
        if isinstance(b, (int, float)) and b == 0:       # sum() starts from 0
            return OpSum([self.operation])
        return OpSum([b.operation if isinstance(b, Gate) else b, self.operation])

    def get_op(self):
        """The reference's per-class matrix builder (gates.py:122-123, 226-232, 303-304 ...): here
        every gate's operator is the symbolic form of its primitive ops."""
        return self.operation

    def set_properties(self):
        """gates.py:125-127, 160-173, 510-512: picks the QuTiP gate function and Pauli of a subclass;
        here those are class attributes (`_kind`, `_axis`), nothing to set."""
        return None

    def set_theta(self, theta):
        return

    def set_phi(self, phi):
        return

    def derivative(self):
        return iden(self.q_N)

    def parameterised_derivative(self, param):
        return self.derivative()

    def flip_pauli(self):
        pass


# %% single-qubit rotations ----------------------------------------------------------------
class PRot(Gate):
    """One-parameter rotation about a Pauli axis on qubit `q_on` (gates.py:106-147)."""

    is_param = True
    param_count = 1
    _kind = None            # primitive opcode
    _axis = None            # 'x' | 'y' | 'z'
    _scale = 1.0            # effective angle = _scale * theta_argument + _offset

    def __init__(self, q_on, q_N):
        self.q_on = q_on
        self.q_N = q_N
        self.theta = 0
        self.phi = 0
        self.is_param = type(self).is_param
        self.param_count = type(self).param_count
        self._sign = 1          # flip_pauli is a no-op on derivatives (quirk Q7)

    @property
    def pauli(self):
        return self._sign * PauliSum.single(self._axis) if self._axis else iden(1)

    @property
    def fock(self):
        return genFockOp(PauliSum.single(self._axis), self.q_on, self.q_N, 2)

    def set_theta(self, theta):
        self.theta = theta

    def _lower(self, slot):
        # self.theta already holds the effective angle; parameters enter with scale/offset
        return [_op(self._kind, self.q_on, param=slot, scale=self._scale,
                    offset=getattr(self, "offset", 0.0))]

    def _frozen(self):
        return [_op(self._kind, self.q_on, offset=self.theta)]

    def derivative(self):
        """-i/2 * Pauli on q_on (gates.py:133-138)."""
        return -1j * self.fock / 2

    def flip_pauli(self):
        self._sign = -self._sign

    def __repr__(self):
        return f"{type(self).__name__}({self.theta:.2f})@q{self.q_on}"


class I(PRot):
    """Identity that still consumes one parameter (quirk Q5, gates.py:150-156)."""
    _kind = _lib.OP_IDENT

    @property
    def fock(self):
        return iden(self.q_N)

    def _frozen(self):
        return [_op(_lib.OP_IDENT, self.q_on)]


class R_x(PRot):
    _kind, _axis = _lib.OP_RX, "x"


class R_y(PRot):
    _kind, _axis = _lib.OP_RY, "y"


class R_z(PRot):
    _kind, _axis = _lib.OP_RZ, "z"


class negative_R_z(R_z):
    """theta -> -theta (gates.py:180-187)."""
    _scale = -1.0

    def set_theta(self, theta):
        self.theta = -1 * theta

    def derivative(self):
        return 1j * self.fock / 2


class offset_R_z(R_z):
    """theta -> theta + offset (gates.py:190-205)."""

    def __init__(self, q_on, q_N, offset):
        super().__init__(q_on, q_N)
        self.offset = offset

    def set_theta(self, theta):
        self.theta = theta + self.offset


# %% fixed single-qubit gates ------------------------------------------------------------------
class H(PRot):
    """Hadamard = x_gate * ry(pi/2) (gates.py:211-232)."""
    is_param = False
    param_count = 0
    _kind = _lib.OP_H
    _fixed_theta = np.pi / 2

    def __init__(self, q_on, q_N):
        super().__init__(q_on, q_N)
        self.theta = self._fixed_theta

    def set_theta(self, angle):
        return None

    def _lower(self, slot):
        return [_op(self._kind, self.q_on)]

    _frozen = lambda self: self._lower(-1)

    def derivative(self):
        raise AttributeError(f"'{type(self).__name__}' object has no attribute 'fock'")


class sqrtH(H):
    def __init__(self, q_on, q_N):
        raise NotImplementedError("sqrtH applies np.sqrt to an operator in the reference "
                                  "(quirk Q6, gates.py:235-242); it is fenced off here")


class X(H):
    _kind = _lib.OP_X


class S(H):
    _kind = _lib.OP_S


class T(H):
    _kind = _lib.OP_T


class fixed_R_y(R_y):
    """gates.py:252-266"""
    is_param = False
    param_count = 0

    def __init__(self, q_on, q_N, theta):
        super().__init__(q_on, q_N)
        self.theta = theta

    def set_theta(self, theta):
        return None

    def _lower(self, slot):
        return [_op(self._kind, self.q_on, offset=self.theta)]


class fixed_R_z(R_z):
    """gates.py:269-283"""
    is_param = False
    param_count = 0

    def __init__(self, q_on, q_N, theta):
        super().__init__(q_on, q_N)
        self.theta = theta

    def set_theta(self, theta):
        return None

    def _lower(self, slot):
        return [_op(self._kind, self.q_on, offset=self.theta)]


# %% entanglers ---------------------------------------------------------------------------------
class EntGate(Gate):
    """Two-qubit fixed gate on (q1, q2) (gates.py:306-324)."""
    _kind = None

    def __init__(self, qs_on, q_N):
        self.q1, self.q2 = qs_on[0], qs_on[1]
        self.q_N = q_N
        self.theta = 0
        self.phi = 0
        self.is_param = False
        self.param_count = 0

    def _lower(self, slot):
        return [] if self._kind is None else [_op(self._kind, self.q1, self.q2)]

    _frozen = lambda self: self._lower(-1)

    def __repr__(self):
        return f"{type(self).__name__}@q{self.q1},q{self.q2}"


class CNOT(EntGate):
    _kind = _lib.OP_CNOT


class CPHASE(EntGate):
    """Defined as a CZ in the reference (quirk Q11, gates.py:333-337)."""
    _kind = _lib.OP_CZ


class sqrtiSWAP(EntGate):
    _kind = _lib.OP_SQRTISWAP


class CZ(EntGate):
    _kind = _lib.OP_CZ


class _Block(EntGate):
    """A fixed block of entanglers applied in `_pairs()` order."""

    def __init__(self, entangler, q_N):
        self.entangler = entangler
        self.q_N = q_N
        self.theta = 0
        self.phi = 0
        self.is_param = False
        self.param_count = 0
        self.members = [entangler(list(pair), q_N) for pair in self._pairs()]

    def _lower(self, slot):
        return flatten([m._lower(-1) for m in self.members])


class CHAIN(_Block):
    """(0,1),(2,3),... then (1,2),(3,4),... in that order (gates.py:355-379)."""

    def _pairs(self):
        N = self.q_N
        return [(2 * j, 2 * j + 1) for j in range(N // 2)] + \
               [(2 * j + 1, 2 * j + 2) for j in range((N - 1) // 2)]

    def __repr__(self):
        return f"CHAIN connected {self.entangler.__name__}s"


class ALLTOALL(_Block):
    """Every ordered pair (quirk Q10, gates.py:382-401)."""

    def _pairs(self):
        return list(permutations(range(self.q_N), 2))

    def __repr__(self):
        return f"ALL connected {self.entangler.__name__}s"


class ARBGATE(Gate):
    """exp(-i theta H) for an arbitrary Hermitian `Ham` on the whole register (gates.py:407-435);
    derivative() = -i H / 2 as in the reference (note: exp(-i theta H), not theta / 2).

    The reference calls a dense Qobj.expm() on every set_theta.  Here H is diagonalised once
    (host LAPACK, H = V diag(lambda) V^dagger) and circuits that contain the gate run as
    segments: gate-program kernels before and after, two dense products with V^dagger / V and a
    diagonal phase for the gate itself (engine.SegmentedProgram, csrc/pqc_dense.cu)."""

    is_param = True
    param_count = 1
    MAX_QUBITS = 13

    def __init__(self, Ham):
        self._Ham = Ham
        mat = np.asarray(Ham.full() if hasattr(Ham, "full") else Ham, dtype=np.complex128)
        D = mat.shape[0]
        n = D.bit_length() - 1
        if mat.ndim != 2 or mat.shape != (D, D) or D != 1 << n:
            raise ValueError("ARBGATE needs a 2^n x 2^n Hamiltonian")
        if n > self.MAX_QUBITS:
            raise NotImplementedError(f"ARBGATE keeps a dense 2^n x 2^n eigenbasis: n <= {self.MAX_QUBITS}")
        if not np.allclose(mat, mat.conj().T, atol=1e-12):
            raise ValueError("ARBGATE needs a Hermitian Hamiltonian")
        super().__init__(n)
        self.theta = 0
        self.pauli = 1
        self._lam, self._V = np.linalg.eigh(mat)

    def _matrix(self, theta):
        return (self._V * np.exp(-1j * theta * self._lam)[None, :]) @ self._V.conj().T

    @property
    def operation(self):
        return qt.DenseOp(self._matrix(self.theta), [[2] * self.q_N, [2] * self.q_N])

    def get_op(self):
        return self.operation

    def set_theta(self, theta):
        self.theta = theta

    def flip_pauli(self):
        self.pauli = -1 * self.pauli            # as in the reference: recorded, never used

    def derivative(self):
        return -1j * self._Ham / 2

    def _lower(self, slot):
        raise NotImplementedError("ARBGATE is not a primitive op: circuits that contain it run "
                                  "through engine.SegmentedProgram")

    def __repr__(self):
        return f"{type(self).__name__}({self.theta:.2f})"


# %% shared parameters ------------------------------------------------------------------------------
def _pauli_commute(a, b):
    (x1, z1), (x2, z2) = a, b
    return (bin(x1 & z2).count("1") + bin(z1 & x2).count("1")) % 2 == 0


class shared_parameter(PRot):
    """One angle drives every member of `layer` (gates.py:441-484)."""

    def __init__(self, layer, q_N, commute=True):
        self.layer = layer
        self.theta = 0
        self.phi = 0
        self.q_N = q_N
        self.is_param = True
        self.param_count = 1
        self.commute = commute

    def set_theta(self, theta):
        self.theta = theta
        for gate in self.layer:
            gate.set_theta(theta)

    def _lower(self, slot):
        # first member acts first: operation = prod(layer[::-1]) (gates.py:475-477)
        return flatten([g._lower(slot) for g in self.layer])

    def _frozen(self):
        return flatten([g._frozen() for g in self.layer])

    def _sum_of_generators_is_exact(self):
        """The reference's commute=False formula (gates.py:458-466) multiplies by the
        element-wise conjugate of the block; that equals the sum-of-generators form iff
        the members' generators mutually commute and every member matrix is symmetric
        (no R_y).  True for the XXZ template's YY+XX blocks (templates.py:248-254)."""
        if self.commute:
            return True
        gens = []
        for g in self.layer:
            if isinstance(g, R_y) or not isinstance(g, (PRot,)) or isinstance(g, (fSim, fixed_fSim)):
                return False
            d = g.derivative()
            gens += list(d.terms)
        return all(_pauli_commute(a, b) for i, a in enumerate(gens) for b in gens[i + 1:])

    def derivative(self):
        if self._sum_of_generators_is_exact():
            deriv = 0
            for g in self.layer:
                deriv = deriv + g.derivative()
            return deriv
        # gates.py:458-466, literally: sum_k prod(layer with member k -> D_k * U_k, reversed) times
        # the ELEMENT-WISE conjugate of the block (take_derivative multiplies the block back in;
        # conj is the inverse only when the block's matrix is symmetric -- reproduced, not fixed).
        # A sum of products of gate programs: an OpSum, applied term by term (PQC's literal
        # derivative path, circuit.py:149-172).
        terms = []
        for count, g in enumerate(self.layer):
            new_layer = list(self.layer)
            new_layer[count] = g.derivative() * g
            terms.append(prod(new_layer[::-1]))
        return OpSum(terms) * self.operation.conj()

    def flip_pauli(self):
        for g in self.layer:
            g.flip_pauli()

    def __repr__(self):
        return f"Block of {self.layer}"


# %% two-qubit Pauli rotations -----------------------------------------------------------------------
class RR(PRot):
    """cos(theta/2) - i sin(theta/2) P(x)P on (q1, q2) (gates.py:492-527)."""

    def __init__(self, qs_on, q_N):
        self.q1, self.q2 = qs_on[0], qs_on[1]
        self.q_N = q_N
        self.theta = 0
        self.phi = 0
        self.is_param = True
        self.param_count = 1
        self._sign = 1

    @property
    def fock1(self):
        return genFockOp(PauliSum.single(self._axis), self.q1, self.q_N, 2)

    @property
    def fock2(self):
        return genFockOp(PauliSum.single(self._axis), self.q2, self.q_N, 2)

    def _lower(self, slot):
        return [_op(self._kind, self.q1, self.q2, param=slot)]

    def _frozen(self):
        return [_op(self._kind, self.q1, self.q2, offset=self.theta)]

    def derivative(self):
        return -1j * (self.fock1 * self.fock2) / 2

    def __repr__(self):
        return f"{type(self).__name__}({self.theta:.2f})@q{self.q1},q{self.q2}"


class R_zz(RR):
    _kind, _axis = _lib.OP_RZZ, "z"


class R_xx(RR):
    _kind, _axis = _lib.OP_RXX, "x"


class R_yy(RR):
    _kind, _axis = _lib.OP_RYY, "y"


class RR_block(shared_parameter):
    """Ring of `rotator` on (i, (i+1) mod N) sharing one angle (gates.py:554-585)."""

    def __init__(self, rotator, q_N):
        self.rotator = rotator
        self.theta = 0
        self.phi = 0
        self.q_N = q_N
        self.is_param = True
        self.param_count = 1
        self.layer = self.gen_layer()
        self.commute = True

    def gen_layer(self):
        N = self.q_N
        return [self.rotator([i, (i + 1) % N], N) for i in range(N)]

    def __repr__(self):
        return f"RR block of {self.layer}"


# %% fSim ----------------------------------------------------------------------------------------------
def _expand_2toN(m4, N, control, target):
    """qutip's gate_expand_2toN on a host matrix: the 4 x 4 `m4` (first factor = `control`) on qubits
    control / target of an N-qubit register, qubit 0 most significant (gates.py:595-597)."""
    if N < 2 or control == target or not (0 <= control < N and 0 <= target < N):
        raise ValueError("control and target must be two different qubits of the register")
    if N > 12:
        raise MemoryError("dense form is only provided for N <= 12")
    D = 1 << N
    b = np.arange(D)
    sc, st = N - 1 - control, N - 1 - target
    col4 = 2 * ((b >> sc) & 1) + ((b >> st) & 1)
    rest = b & ~((1 << sc) | (1 << st))
    out = np.zeros((D, D), dtype=np.complex128)
    for r4 in range(4):
        rows = rest | ((r4 >> 1) << sc) | ((r4 & 1) << st)
        out[rows, b] = m4[r4, col4]
    return out


def _fsim_matrix(theta, phi, which):
    """The 4 x 4 matrices of gates.py:598-606 (which = 0), 619-627 (1: "d/dtheta") and 640-648
    (2: "d/dphi") exactly as written there -- the |00> entry of both derivative matrices stays 1 and
    d/dtheta keeps exp(-i phi) (quirk Q3 of SURVEY.md)."""
    c, s, e = np.cos(theta), np.sin(theta), np.exp(-1j * phi)
    m = np.zeros((4, 4), dtype=np.complex128)
    m[0, 0] = 1
    if which == 1:
        m[1, 1] = m[2, 2] = -s
        m[1, 2] = m[2, 1] = -1j * c
    else:
        m[1, 1] = m[2, 2] = c
        m[1, 2] = m[2, 1] = -1j * s
    m[3, 3] = -1j * e if which == 2 else e
    return m


def _two_qubit_dense(m4, N, control, target):
    if (control == 1 and target == 0) and N is None:          # gates.py:591-592
        N = 2
    if N is None:
        return qt.DenseOp(m4, [[2, 2], [2, 2]])
    return qt.DenseOp(_expand_2toN(m4, N, control, target), [[2] * N, [2] * N])


def fsim_gate(theta, phi, N=None, control=0, target=1):
    """gates.py:588-606.  On a register (N given) the gate is the symbolic one-op Operator -- it acts
    on device states at any N and `.full()` gives the matrix; without N the bare 4 x 4 matrix."""
    if (control == 1 and target == 0) and N is None:
        N = 2
    if N is None:
        return qt.DenseOp(_fsim_matrix(theta, phi, 0), [[2, 2], [2, 2]])
    return Operator(N, [_op(_lib.OP_FSIM, control, target, scale=phi, offset=theta)])


def fsim_gate_d_theta(theta, phi, N=None, control=0, target=1):
    """gates.py:609-627 (dense: N <= 12; derivative STATES at any N come from PQC.get_gradients)."""
    return _two_qubit_dense(_fsim_matrix(theta, phi, 1), N, control, target)


def fsim_gate_d_phi(theta, phi, N=None, control=0, target=1):
    """gates.py:630-648"""
    return _two_qubit_dense(_fsim_matrix(theta, phi, 2), N, control, target)


def fixed_fsim_gate(theta, N=None, control=0, target=1):
    """gates.py:700-716: fSim with phi = 0."""
    if (control == 1 and target == 0) and N is None:
        N = 2
    if N is None:
        return qt.DenseOp(_fsim_matrix(theta, 0.0, 0), [[2, 2], [2, 2]])
    return Operator(N, [_op(_lib.OP_FIXED_FSIM, control, target, offset=theta)])


def fixed_fsim_gate_d_theta(theta, N=None, control=0, target=1):
    """gates.py:719-737"""
    return _two_qubit_dense(_fsim_matrix(theta, 0.0, 1), N, control, target)


class fSim(PRot):
    """Two-parameter fSim(theta, phi) (gates.py:588-606,651-697)."""

    def __init__(self, qs_on, q_N):
        self.q1, self.q2 = qs_on
        self.q_N = q_N
        self.theta = 0
        self.phi = 0
        self.is_param = True
        self.param_count = 2

    def set_phi(self, phi):
        self.phi = phi

    def _lower(self, slot):
        return [_op(_lib.OP_FSIM, self.q1, self.q2, param=slot,
                    param2=slot + 1 if slot >= 0 else -1)]

    def _frozen(self):
        # fixed fSim: theta rides in `offset`, phi in `scale` (include/pqc_b200.h)
        return [_op(_lib.OP_FSIM, self.q1, self.q2, scale=self.phi, offset=self.theta)]

    def derivative(self):
        # the reference's fSim inherits PRot.derivative (gates.py:133-139), which reads a `fock`
        # attribute fSim.__init__ never sets (gates.py:652-660): an AttributeError there as well
        raise AttributeError("fSim has no single derivative (no `fock` generator, as in the reference); "
                             "use parameterised_derivative(1 | 2)")

    def parameterised_derivative(self, param):
        """gates.py:680-693: param 1 -> 'd/dtheta', 2 -> 'd/dphi' as dense operators (N <= 12; the
        derivative STATES of PQC.get_gradients / take_derivative come from the op's kernels)."""
        if param == 1:
            return fsim_gate_d_theta(self.theta, self.phi, N=self.q_N, control=self.q1, target=self.q2)
        if param == 2:
            return fsim_gate_d_phi(self.theta, self.phi, N=self.q_N, control=self.q1, target=self.q2)
        raise ValueError("fSim has parameters 1 (theta) and 2 (phi)")

    def flip_pauli(self):
        pass

    def __repr__(self):
        return f"{type(self).__name__}({self.theta:.2f},{self.phi:.2f})@q{self.q1, self.q2}"


class fixed_fSim(PRot):
    """fSim with phi = 0 (gates.py:700-759)."""

    def __init__(self, qs_on, q_N):
        self.q1, self.q2 = qs_on
        self.q_N = q_N
        self.theta = 0
        self.phi = 0
        self.is_param = True
        self.param_count = 1

    def _lower(self, slot):
        return [_op(_lib.OP_FIXED_FSIM, self.q1, self.q2, param=slot)]

    def _frozen(self):
        return [_op(_lib.OP_FIXED_FSIM, self.q1, self.q2, offset=self.theta)]

    def derivative(self):
        """gates.py:753-756 as a dense operator (N <= 12; derivative STATES at any N come from
        PQC.get_gradients / take_derivative)."""
        return fixed_fsim_gate_d_theta(self.theta, N=self.q_N, control=self.q1, target=self.q2)

    def flip_pauli(self):
        pass
