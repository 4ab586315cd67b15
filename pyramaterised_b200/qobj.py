"""Host-side stand-ins for the `qutip.Qobj` subset the reference touches.

The reference keeps every state and operator as a QuTiP ``Qobj``.  Here
  * ``State``     -- a ket living in device memory; exposes what the reference calls on
                     states: overlap, ptrace(k), tr, dims, data.toarray(), ==, [i][0][0]
                     (measure.py:52,58,135,233-236,335-336; tests.py:71-72,86);
  * ``PauliSum``  -- a symbolic sum of Pauli strings (Hamiltonians, gate generators):
                     what genFockOp / sigmaz products evaluate to (gates.py:39-46,
                     circuit.py:28-31, templates.py:220-227);
  * ``Operator``  -- a frozen sequence of primitive gate ops, what ``Gate.operation``
                     stands for (gates.py:57) without ever building a 2^n x 2^n matrix;
  * ``DenseOp``   -- tiny host matrices (the 2x2 result of ptrace).
All state arithmetic runs in the CUDA library through ``engine``.
"""
import numbers

import numpy as np
import torch

from . import _lib, engine

ATOL = 1e-12          # qutip.settings.atol, used by Qobj.__eq__ (tests.py:86)


def _is_scalar(x):
    return isinstance(x, (numbers.Number, np.number))


# ---------------------------------------------------------------------------------------
class _DataView:
    """state.data.toarray() -> column vector (measure.py:335)."""

    def __init__(self, owner):
        self._o = owner

    def toarray(self):
        return self._o.full()

    @property
    def shape(self):
        return self._o.shape


class DenseOp:
    """Small host matrix with the Qobj operator surface used on ptrace results
    (measure.py:233-235: `rho *= rho; rho.tr()`)."""

    def __init__(self, mat, dims=None):
        self._m = np.asarray(mat, dtype=np.complex128)
        k = self._m.shape[0].bit_length() - 1
        self.dims = dims or [[2] * k, [2] * k]

    type = "oper"

    @property
    def shape(self):
        return self._m.shape

    def full(self):
        return self._m.copy()

    @property
    def data(self):
        return _DataView(self)

    def __mul__(self, o):
        if isinstance(o, DenseOp):
            return DenseOp(self._m @ o._m, self.dims)
        if _is_scalar(o):
            return DenseOp(self._m * o, self.dims)
        if isinstance(o, State):
            # dense operator on a ket (what `ARBGATE * state` is in the reference, gates.py:63-67):
            # one row-major matrix-vector product in the library (csrc/pqc_dense.cu)
            if self._m.shape[1] != o.tensor.numel():
                raise TypeError("incompatible dimensions")
            M = torch.from_numpy(np.ascontiguousarray(self._m)).to(o.tensor.device)
            return State(engine.dense_apply(o.tensor.reshape(1, -1), M)[0], o.dims)
        return NotImplemented

    __rmul__ = lambda self, o: DenseOp(self._m * o, self.dims) if _is_scalar(o) else NotImplemented

    def __truediv__(self, o):
        return DenseOp(self._m / o, self.dims) if _is_scalar(o) else NotImplemented

    def __add__(self, o):
        if isinstance(o, DenseOp):
            return DenseOp(self._m + o._m, self.dims)
        if _is_scalar(o):
            return self if o == 0 else DenseOp(self._m + o * np.eye(len(self._m)), self.dims)
        return NotImplemented

    __radd__ = __add__

    def dag(self):
        return DenseOp(self._m.conj().T, self.dims)

    @property
    def isherm(self):
        return bool(np.all(np.abs(self._m - self._m.conj().T) < ATOL))

    def tr(self):
        t = np.trace(self._m)
        return float(t.real) if self.isherm else complex(t)

    def __eq__(self, o):
        return isinstance(o, DenseOp) and self._m.shape == o._m.shape and \
            bool(np.all(np.abs(self._m - o._m) < ATOL))

    __hash__ = None

    def __repr__(self):
        return f"DenseOp(dims={self.dims})\n{self._m}"


# ---------------------------------------------------------------------------------------
class State:
    """A ket in device memory (complex128 [D])."""

    type = "ket"
    isket = True

    def __init__(self, data, dims=None):
        if isinstance(data, State):
            t, dims = data._t.clone(), dims or data.dims
        elif isinstance(data, torch.Tensor):
            t = data
            if t.dtype != torch.complex128 or not t.is_cuda:
                t = engine.as_states(t)
        else:
            a = np.asarray(data, dtype=np.complex128).reshape(-1)
            t = engine.as_states(a)
        self._t = t.reshape(-1)
        D = self._t.numel()
        if dims is None:
            n = D.bit_length() - 1
            dims = [[2] * n, [1] * n] if (1 << n) == D else [[D], [1]]
        self.dims = [list(dims[0]), list(dims[1])]

    # ---- views ------------------------------------------------------------------------
    @property
    def tensor(self):
        return self._t

    @property
    def n_qubits(self):
        return len(self.dims[0])

    @property
    def shape(self):
        return (self._t.numel(), 1)

    def numpy(self):
        return self._t.cpu().numpy()

    def full(self):
        return self.numpy().reshape(-1, 1)

    @property
    def data(self):
        return _DataView(self)

    def __getitem__(self, ind):
        """Qobj[i] -> [[amp]] so that out[i][0][0] reads one amplitude (tests.py:71-72)."""
        v = self._t[ind]
        return v.cpu().numpy().reshape(-1, 1) if v.dim() else np.array([[v.item()]])

    def copy(self):
        return State(self._t.clone(), self.dims)

    # ---- the calls the reference makes ----------------------------------------------------
    def overlap(self, other):
        """<self|other> (measure.py:52,58,135)."""
        if not isinstance(other, State):
            other = State(other)
        return complex(engine.overlap(self._t, other._t)[0].item())

    def ptrace(self, sel):
        """Reduced density matrix of ONE qubit (measure.py:233)."""
        if isinstance(sel, (list, tuple)):
            if len(sel) != 1:
                raise NotImplementedError("ptrace is provided for a single qubit (single_Q)")
            sel = sel[0]
        rho = engine.ptrace_1q(self._t, int(sel)).cpu().numpy()
        return DenseOp(rho, [[2], [2]])

    def norm(self):
        return float(np.sqrt(abs(self.overlap(self))))

    def unit(self):
        return self * (1.0 / self.norm())

    def dag(self):
        raise NotImplementedError("bras are not materialised; use State.overlap")

    def __eq__(self, other):
        if not isinstance(other, State) or self.dims != other.dims:
            return False
        return bool(np.all(np.abs(self.numpy() - other.numpy()) < ATOL))

    __hash__ = None

    # ---- light arithmetic (state preparation only; never on the hot path) -------------------
    def __mul__(self, o):
        if _is_scalar(o):
            return State(self._t * complex(o), self.dims)
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self * (1.0 / o)

    def __neg__(self):
        return self * -1.0

    def __add__(self, o):
        if isinstance(o, State):
            return State(self._t + o._t, self.dims)
        if _is_scalar(o) and o == 0:
            return self
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, o):
        return self + (-1.0 * o)

    def __repr__(self):
        return f"State(dims={self.dims}, shape={self.shape}, type=ket, device={self._t.device})"


def basis(N, k=0):
    v = np.zeros(N, dtype=np.complex128)
    v[k] = 1
    return State(v, [[N], [1]])


def tensor(*args):
    """Kronecker product, first factor most significant (circuit.py:22)."""
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    if all(isinstance(a, PauliSum) for a in args):
        return PauliSum.tensor(args)
    if not all(isinstance(a, State) for a in args):
        raise TypeError("tensor() takes States or PauliSums")
    out = np.array([1.0 + 0j])
    d0, d1 = [], []
    for a in args:
        out = np.kron(out, a.numpy())
        d0 += a.dims[0]
        d1 += a.dims[1]
    return State(out, [d0, d1])


def basis_state(n, index):
    """|index> on n qubits without a host-side 2^n array round trip for large n."""
    t = torch.zeros(1 << n, dtype=torch.complex128, device=engine.device())
    t[index] = 1.0
    return State(t, [[2] * n, [1] * n])


# ---------------------------------------------------------------------------------------
class PauliSum:
    """sum_t coef_t * P_t on n qubits; P_t = prod_q i^{x_q z_q} X^{x_q} Z^{z_q}
    (so x&z marks a Y).  Masks here are indexed by QUBIT (bit q of the mask = qubit q)."""

    type = "oper"

    def __init__(self, n, terms=None):
        self.n = int(n)
        self.terms = {}
        for (x, z), c in (terms or {}).items():
            if c != 0:
                self.terms[(int(x), int(z))] = complex(c)

    # ---- constructors -------------------------------------------------------------------
    @staticmethod
    def identity(n):
        return PauliSum(n, {(0, 0): 1.0})

    @staticmethod
    def single(kind, n=1, q=0):
        x = 1 << q if kind in "xy" else 0
        z = 1 << q if kind in "yz" else 0
        return PauliSum(n, {(x, z): 1.0})

    @staticmethod
    def tensor(factors):
        out = PauliSum(0, {(0, 0): 1.0})
        for f in factors:
            new = {}
            for (x1, z1), c1 in out.terms.items():
                for (x2, z2), c2 in f.terms.items():
                    key = (x1 | (x2 << out.n), z1 | (z2 << out.n))
                    new[key] = new.get(key, 0) + c1 * c2
            out = PauliSum(out.n + f.n, new)
        return out

    @property
    def dims(self):
        return [[2] * self.n, [2] * self.n]

    @property
    def shape(self):
        return (1 << self.n, 1 << self.n)

    @property
    def isherm(self):
        return all(abs(c.imag) < ATOL for c in self.terms.values())

    # ---- algebra ---------------------------------------------------------------------------
    def _lift(self, o):
        if _is_scalar(o):
            return PauliSum(self.n, {(0, 0): o})
        return o

    def __add__(self, o):
        if _is_scalar(o) and o == 0:
            return self
        o = self._lift(o)
        if not isinstance(o, PauliSum):
            return NotImplemented
        if o.n != self.n:
            raise TypeError("Incompatible quantum object dimensions")
        t = dict(self.terms)
        for k, c in o.terms.items():
            t[k] = t.get(k, 0) + c
        return PauliSum(self.n, t)

    __radd__ = __add__

    def __neg__(self):
        return self * -1.0

    def __sub__(self, o):
        return self + (-1.0 * self._lift(o))

    def __rsub__(self, o):
        return (-self) + o

    def __truediv__(self, o):
        return self * (1.0 / o)

    def __mul__(self, o):
        if _is_scalar(o):
            return PauliSum(self.n, {k: c * o for k, c in self.terms.items()})
        if isinstance(o, PauliSum):
            if o.n != self.n:
                raise TypeError("Incompatible Qobj shapes")
            t = {}
            for (x1, z1), c1 in self.terms.items():
                for (x2, z2), c2 in o.terms.items():
                    x3, z3 = x1 ^ x2, z1 ^ z2
                    k = bin(x1 & z1).count("1") + bin(x2 & z2).count("1") - bin(x3 & z3).count("1") \
                        + 2 * bin(z1 & x2).count("1")
                    t[(x3, z3)] = t.get((x3, z3), 0) + c1 * c2 * (1j ** (k % 4))
            return PauliSum(self.n, t)
        if isinstance(o, State):
            out = engine.pauli_apply(o.tensor.reshape(1, -1), self.device_terms())
            return State(out[0], o.dims)
        if isinstance(o, (Operator, Composite)):
            return Composite([self]) * o
        return NotImplemented        # e.g. PauliSum * Gate -> Gate.__rmul__ (gates.py:69-73)

    def __rmul__(self, o):
        if _is_scalar(o):
            return self * o
        return NotImplemented

    def conj(self):
        """element-wise conjugate: Y -> -Y, coefficients conjugated."""
        return PauliSum(self.n, {k: np.conj(c) * (-1) ** bin(k[0] & k[1]).count("1")
                                 for k, c in self.terms.items()})

    def dag(self):
        return PauliSum(self.n, {k: np.conj(c) for k, c in self.terms.items()})

    # ---- evaluation ----------------------------------------------------------------------
    def device_terms(self):
        """[(xmask, zmask, coef)] with masks in basis-index bit positions (bit b = qubit n-1-b)."""
        out = []
        for (x, z), c in self.terms.items():
            xb = sum(1 << (self.n - 1 - q) for q in range(self.n) if x >> q & 1)
            zb = sum(1 << (self.n - 1 - q) for q in range(self.n) if z >> q & 1)
            out.append((xb, zb, c))
        return out

    def expect(self, state):
        v = complex(engine.pauli_expect(state.tensor.reshape(1, -1), self.device_terms())[0].item())
        return float(v.real) if self.isherm else v

    def full(self):
        """Dense host matrix -- inspection / small-n eigen-decomposition only."""
        if self.n > 12:
            raise MemoryError("dense form is only provided for n <= 12")
        D = 1 << self.n
        idx = np.arange(D)
        M = np.zeros((D, D), dtype=np.complex128)
        for xb, zb, c in self.device_terms():
            ny = bin(xb & zb).count("1")
            par = np.zeros(D, dtype=np.int64)
            m = idx & zb
            for b in range(self.n):
                par ^= (m >> b) & 1
            M[idx ^ xb, idx] += c * (1j ** ny) * (1 - 2 * par)
        return M

    def eigenenergies(self):
        return np.linalg.eigvalsh(self.full())

    def groundstate(self):
        w, v = np.linalg.eigh(self.full())
        return w[0], State(v[:, 0], [[2] * self.n, [1] * self.n])

    def __eq__(self, o):
        if not isinstance(o, PauliSum) or o.n != self.n:
            return False
        keys = set(self.terms) | set(o.terms)
        return all(abs(self.terms.get(k, 0) - o.terms.get(k, 0)) < ATOL for k in keys)

    __hash__ = None

    def __repr__(self):
        def name(x, z):
            return "".join("IXZY"[(x >> q & 1) + 2 * (z >> q & 1)] for q in range(self.n))
        return "PauliSum(" + " + ".join(f"({c:.4g})*{name(x, z)}"
                                         for (x, z), c in self.terms.items()) + ")"


def qeye(n):
    if isinstance(n, (list, tuple)):
        return PauliSum.identity(len(n))
    if n != 2:
        raise NotImplementedError("only qubit identities are provided")
    return PauliSum.identity(1)


def sigmax():
    return PauliSum.single("x")


def sigmay():
    return PauliSum.single("y")


def sigmaz():
    return PauliSum.single("z")


def expect(oper, state):
    """qt.expect(H, psi) (circuit.py:136)."""
    return oper.expect(state)


# ---------------------------------------------------------------------------------------
_CONJ_NEGATES = (_lib.OP_RX, _lib.OP_RZ, _lib.OP_RXX, _lib.OP_RYY, _lib.OP_RZZ)


class Operator:
    """A frozen product of primitive ops (all angles fixed): what `Gate.operation`
    denotes.  ops are applied in list order."""

    type = "oper"

    def __init__(self, n, ops):
        self.n = int(n)
        self.ops = [tuple(o) for o in ops]

    @property
    def dims(self):
        return [[2] * self.n, [2] * self.n]

    def _apply(self, t):
        """t: [S, D] device tensor -> [S, D]"""
        prog = engine.Program(self.n, 0, self.ops)
        return prog.run(None, init=t)

    def __mul__(self, o):
        if isinstance(o, State):
            return State(self._apply(o.tensor.reshape(1, -1))[0], o.dims)
        if isinstance(o, Operator):
            return Operator(self.n, o.ops + self.ops)
        if isinstance(o, (PauliSum, Composite)):
            return Composite([self]) * o
        return NotImplemented        # Operator * Gate -> Gate.__rmul__

    def __rmul__(self, o):
        if _is_scalar(o) and o == 1:      # prod() starts from the integer 1 (gates.py:30-31)
            return self
        return NotImplemented

    def conj(self):
        """element-wise complex conjugate of the matrix (Qobj.conj, gates.py:465)."""
        out = []
        for (kind, q0, q1, p, p2, g, scale, offset) in self.ops:
            if kind in _CONJ_NEGATES:
                out.append((kind, q0, q1, p, p2, g, -scale, -offset))
            elif kind == _lib.OP_S:
                out += [(kind, q0, q1, p, p2, g, scale, offset)] * 3
            elif kind == _lib.OP_T or kind == _lib.OP_SQRTISWAP:
                out += [(kind, q0, q1, p, p2, g, scale, offset)] * 7
            elif kind in (_lib.OP_FSIM, _lib.OP_FIXED_FSIM):
                raise NotImplementedError("conj() of fSim operators")
            else:
                out.append((kind, q0, q1, p, p2, g, scale, offset))
        return Operator(self.n, out)

    def full(self):
        """Dense matrix by applying the ops to every basis state on the device (n <= 12)."""
        if self.n > 12:
            raise MemoryError("dense form is only provided for n <= 12")
        D = 1 << self.n
        eye = torch.eye(D, dtype=torch.complex128, device=engine.device())
        return self._apply(eye).cpu().numpy().T

    def __repr__(self):
        return f"Operator(n={self.n}, {[_lib.OP_NAMES[o[0]] for o in self.ops]})"


class OpSum:
    """Sum of operator-like terms (Operator / PauliSum / Composite / DenseOp): what
    `Gate.__add__` / `__radd__` denote (gates.py:75-85, `self.operation + b.operation`).  Unitaries
    are kept as gate programs, so their sum stays a list of terms: applied to a state term by
    term, dense (`full()`) only on request."""

    type = "oper"

    def __init__(self, terms):
        self.terms = []
        for t in terms:
            self.terms += t.terms if isinstance(t, OpSum) else [t]

    @property
    def dims(self):
        return self.terms[0].dims

    def __add__(self, o):
        if _is_scalar(o) and o == 0:          # sum() starts from the integer 0
            return self
        if hasattr(o, "operation") and not isinstance(o, (State, OpSum)):
            o = o.operation
        if isinstance(o, (OpSum, Operator, PauliSum, Composite, DenseOp)):
            return OpSum([self, o])
        return NotImplemented

    __radd__ = lambda self, o: self if (_is_scalar(o) and o == 0) else OpSum([o, self]) \
        if isinstance(o, (Operator, PauliSum, Composite, DenseOp)) else NotImplemented

    def __mul__(self, o):
        if isinstance(o, State):
            out = None
            for t in self.terms:
                if isinstance(t, DenseOp):
                    m = torch.as_tensor(t._m, device=o.tensor.device)
                    v = State(m @ o.tensor.reshape(-1), o.dims)
                else:
                    v = t * o
                out = v if out is None else out + v
            return out
        if isinstance(o, (Operator, PauliSum, Composite)):
            return OpSum([Composite([t]) * o if not isinstance(t, Composite) else t * o
                          for t in self.terms])
        return NotImplemented

    def full(self):
        return sum(t.full() for t in self.terms)

    def __repr__(self):
        return "OpSum(" + " + ".join(repr(t) for t in self.terms) + ")"


class Composite:
    """Product of Operator / PauliSum factors, leftmost acts last."""

    type = "oper"

    def __init__(self, factors):
        self.factors = list(factors)

    def __mul__(self, o):
        if isinstance(o, State):
            for f in reversed(self.factors):
                o = f * o
            return o
        if isinstance(o, Composite):
            return Composite(self.factors + o.factors)
        if isinstance(o, (Operator, PauliSum)):
            return Composite(self.factors + [o])
        return NotImplemented

    def __rmul__(self, o):
        if _is_scalar(o) and o == 1:
            return self
        return NotImplemented

    def full(self):
        """Dense matrix (small n): the product of the factors' matrices."""
        out = None
        for f in self.factors:
            m = f.full()
            out = m if out is None else out @ m
        return out
