"""qutip_lite -- a CPU restatement of the QuTiP 4.7.2 subset that the reference calls.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyramaterised_b200/`` may import this
module; it exists so that

  * ``tests/golden/make_golden.py`` can import the *unmodified* reference package
    from ``/root/reference`` (whose only numerical backend, ``qutip==4.7.2``
    -- /root/reference/requirements.txt:15 -- is not installable offline) and
    record golden vectors produced by the reference's own Python code, and
  * the CPU test-suite can cross check ``oracle/pqc_oracle.py`` against a dense,
    literal "every gate is a 2^n x 2^n operator" execution.

QuTiP's source is not vendored under /root/reference, so the semantics below are
restated from the published QuTiP 4.7 behaviour and are
pinned by the reference's own known-answer tests (tests.py:192-212 golden energy
and EQD, tests.py:114-128 NPQC QFIM = identity, tests.py:284-295 Bell state,
tests.py:64-87 ry(pi/2) and H.H) -- see tests/test_oracle_golden.py.

Call sites in the reference that define the required surface:
  gates.py:39-46 (qeye, tensor), gates.py:63-85 (Qobj * and +),
  gates.py:122-123,161-174 (qip.operations.rx/ry/rz), gates.py:228-249 (x_gate),
  gates.py:289-300 (phasegate, t_gate), gates.py:329-349 (cnot, cz_gate,
  sqrtiswap), gates.py:417 (expm), gates.py:465 (conj), gates.py:533-550
  (x/y/z_gate), gates.py:595-597 (gate_expand_2toN), circuit.py:22,29-31,136,141
  (basis, sigmaz, expect, overlap), measure.py:52,135,233-235,335-336 (overlap,
  ptrace, tr, data.toarray, dims), tests.py:26,19,45 (bell_state, rand_ket_haar,
  rand_unitary_haar), tests.py:140-141 (groundstate, eigenenergies).
"""
from __future__ import annotations

import numbers
import sys
import types

import numpy as np
import scipy.linalg
import scipy.sparse as sp

__version__ = "4.7.2-lite"


class _Settings:
    auto_tidyup = True
    auto_tidyup_atol = 1e-12
    atol = 1e-12


settings = _Settings()


def _csr(a):
    m = sp.csr_matrix(a, dtype=np.complex128)
    m.sort_indices()
    return m


def _tidy(m):
    """QuTiP auto-tidyup: real and imaginary parts below atol are zeroed
    independently, then explicit zeros are dropped."""
    if not settings.auto_tidyup:
        return m
    d = m.data
    if d.size:
        re = np.where(np.abs(d.real) < settings.auto_tidyup_atol, 0.0, d.real)
        im = np.where(np.abs(d.imag) < settings.auto_tidyup_atol, 0.0, d.imag)
        m.data = re + 1j * im
        m.eliminate_zeros()
    return m


class Qobj:
    __array_priority__ = 100

    def __init__(self, inpt=None, dims=None, shape=None, **_ignored):
        if isinstance(inpt, Qobj):
            self.data = inpt.data.copy()
            self.dims = [list(inpt.dims[0]), list(inpt.dims[1])] if dims is None else dims
            return
        if inpt is None:
            inpt = [[0]]
        if sp.issparse(inpt):
            self.data = _csr(inpt)
        else:
            arr = np.asarray(inpt, dtype=np.complex128)
            if arr.ndim == 0:
                arr = arr.reshape(1, 1)
            elif arr.ndim == 1:
                arr = arr.reshape(-1, 1)      # 1-d input is a ket
            self.data = _csr(arr)
        if dims is None:
            dims = [[self.data.shape[0]], [self.data.shape[1]]]
        self.dims = [list(dims[0]), list(dims[1])]

    # ----- structure -----------------------------------------------------------
    @property
    def shape(self):
        return self.data.shape

    @property
    def type(self):
        r, c = self.shape
        if c == 1 and r != 1:
            return "ket"
        if r == 1 and c != 1:
            return "bra"
        return "oper"

    @property
    def isket(self):
        return self.type == "ket"

    @property
    def isherm(self):
        if self.shape[0] != self.shape[1]:
            return False
        d = (self.data - self.data.getH())
        return d.nnz == 0 or np.max(np.abs(d.data)) < settings.atol

    def full(self):
        return self.data.toarray()

    def __array__(self, dtype=None, copy=None):
        return self.full() if dtype is None else self.full().astype(dtype)

    def copy(self):
        return Qobj(self)

    # ----- algebra ---------------------------------------------------------------
    @staticmethod
    def _new(data, dims):
        out = Qobj.__new__(Qobj)
        out.data = _tidy(_csr(data))
        out.dims = [list(dims[0]), list(dims[1])]
        return out

    def __mul__(self, other):
        if isinstance(other, Qobj):
            if self.shape[1] != other.shape[0]:
                raise TypeError("Incompatible Qobj shapes")
            return Qobj._new(self.data @ other.data, [self.dims[0], other.dims[1]])
        if isinstance(other, (numbers.Number, np.number)):
            return Qobj._new(self.data * complex(other), self.dims)
        return NotImplemented          # lets Gate.__rmul__ take over (gates.py:69-73)

    def __rmul__(self, other):
        if isinstance(other, (numbers.Number, np.number)):
            return Qobj._new(self.data * complex(other), self.dims)
        return NotImplemented

    def __truediv__(self, other):
        if isinstance(other, (numbers.Number, np.number)):
            return Qobj._new(self.data / complex(other), self.dims)
        return NotImplemented

    def __neg__(self):
        return Qobj._new(-self.data, self.dims)

    def __add__(self, other):
        if not isinstance(other, Qobj):
            if isinstance(other, (numbers.Number, np.number)):
                other = Qobj(other)
            else:
                return NotImplemented
        if other.shape == (1, 1) and self.shape != (1, 1):
            c = other.data[0, 0] if other.data.nnz else 0.0
            if c == 0:
                return self
            if self.type == "oper":
                return Qobj._new(self.data + c * sp.identity(self.shape[0], format="csr"),
                                 self.dims)
            dat = self.data.copy()
            dat.data = dat.data + c
            return Qobj._new(dat, self.dims)
        if self.shape == (1, 1) and other.shape != (1, 1):
            return other.__add__(self)
        if self.dims != other.dims:
            raise TypeError("Incompatible quantum object dimensions")
        return Qobj._new(self.data + other.data, self.dims)

    def __radd__(self, other):
        return self + other

    def __sub__(self, other):
        return self + (-1 * other)

    def __rsub__(self, other):
        return (-self) + other

    def __eq__(self, other):
        if not isinstance(other, Qobj) or self.dims != other.dims:
            return False
        d = self.data - other.data
        return d.nnz == 0 or bool(np.all(np.abs(d.data) < settings.atol))

    __hash__ = None

    def __getitem__(self, ind):
        out = self.data[ind]
        return out.toarray() if sp.issparse(out) else out

    def dag(self):
        return Qobj._new(self.data.getH(), [self.dims[1], self.dims[0]])

    def conj(self):
        return Qobj._new(self.data.conj(), self.dims)

    def trans(self):
        return Qobj._new(self.data.T, [self.dims[1], self.dims[0]])

    def tr(self):
        t = self.data.diagonal().sum()
        return float(t.real) if self.isherm else complex(t)

    def norm(self):
        if self.type in ("ket", "bra"):
            return float(np.sqrt(np.sum(np.abs(self.data.data) ** 2)))
        return float(np.sum(scipy.linalg.svdvals(self.full())))

    def unit(self):
        return self / self.norm()

    def expm(self):
        return Qobj._new(scipy.linalg.expm(self.full()), self.dims)

    def overlap(self, other):
        """<self|other> for two kets (measure.py:52,58,135)."""
        if self.type == "ket" and other.type == "ket":
            return complex((self.data.getH() @ other.data).toarray()[0, 0])
        if self.type == "bra" and other.type == "ket":
            return complex((self.data @ other.data).toarray()[0, 0])
        if self.type == "ket" and other.type == "bra":
            return complex((other.data @ self.data).toarray()[0, 0].conjugate())
        raise TypeError("overlap restated for kets/bras only")

    def ptrace(self, sel):
        """Reduced density matrix on the selected subsystems (measure.py:233)."""
        if isinstance(sel, numbers.Integral):
            sel = [int(sel)]
        sel = sorted(sel)
        dims = self.dims[0]
        n = len(dims)
        if self.type == "ket":
            psi = self.full().reshape(dims)
            rest = [k for k in range(n) if k not in sel]
            m = np.transpose(psi, sel + rest).reshape(
                int(np.prod([dims[k] for k in sel])), -1)
            rho = m @ m.conj().T
        else:
            rho_full = self.full().reshape(dims + dims)
            rest = [k for k in range(n) if k not in sel]
            perm = sel + rest + [n + k for k in sel] + [n + k for k in rest]
            ds = int(np.prod([dims[k] for k in sel]))
            dr = int(np.prod([dims[k] for k in rest])) if rest else 1
            r = np.transpose(rho_full, perm).reshape(ds, dr, ds, dr)
            rho = np.einsum("arbr->ab", r)
        sd = [dims[k] for k in sel]
        return Qobj._new(rho, [sd, sd])

    def permute(self, order):
        dims = self.dims[0]
        n = len(dims)
        if self.type == "ket":
            a = self.full().reshape(dims)
            a = np.transpose(a, order)
            nd = [dims[k] for k in order]
            return Qobj._new(a.reshape(-1, 1), [nd, [1] * n])
        a = self.full().reshape(dims + dims)
        a = np.transpose(a, list(order) + [n + k for k in order])
        nd = [dims[k] for k in order]
        return Qobj._new(a.reshape(self.shape), [nd, nd])

    def eigenenergies(self):
        return np.linalg.eigvalsh(self.full())

    def eigenstates(self):
        w, v = np.linalg.eigh(self.full())
        kets = [Qobj._new(v[:, k].reshape(-1, 1), [self.dims[0], [1] * len(self.dims[0])])
                for k in range(len(w))]
        return w, kets

    def groundstate(self):
        w, kets = self.eigenstates()
        return w[0], kets[0]

    def __repr__(self):
        return f"Qobj(dims={self.dims}, shape={self.shape}, type={self.type})\n{self.full()}"


# ----- states / operators -------------------------------------------------------

def basis(N, n=0):
    v = np.zeros((N, 1), dtype=np.complex128)
    v[n, 0] = 1.0
    return Qobj(v, dims=[[N], [1]])


def qeye(N):
    if isinstance(N, (list, tuple)):
        d = int(np.prod(N))
        return Qobj(sp.identity(d, format="csr", dtype=np.complex128), dims=[list(N), list(N)])
    return Qobj(sp.identity(int(N), format="csr", dtype=np.complex128), dims=[[int(N)], [int(N)]])


identity = qeye


def sigmax():
    return Qobj([[0, 1], [1, 0]], dims=[[2], [2]])


def sigmay():
    return Qobj([[0, -1j], [1j, 0]], dims=[[2], [2]])


def sigmaz():
    return Qobj([[1, 0], [0, -1]], dims=[[2], [2]])


def tensor(*args):
    """Kronecker product, first factor most significant (gates.py:42, circuit.py:22)."""
    if len(args) == 1 and isinstance(args[0], (list, tuple, np.ndarray)):
        args = tuple(args[0])
    out = None
    d0, d1 = [], []
    for q in args:
        out = q.data if out is None else sp.kron(out, q.data, format="csr")
        d0 += list(q.dims[0])
        d1 += list(q.dims[1])
    return Qobj._new(out, [d0, d1])


def expect(oper, state):
    """<psi|O|psi> for a ket (circuit.py:136); real for Hermitian O."""
    if isinstance(state, (list, tuple)):
        return np.array([expect(oper, s) for s in state])
    if state.type == "ket":
        v = (state.data.getH() @ (oper.data @ state.data)).toarray()[0, 0]
    else:
        v = (oper.data @ state.data).diagonal().sum()
    return float(v.real) if oper.isherm else complex(v)


def bell_state(state="00"):
    """bell_state('11') = (|01> - |10>)/sqrt(2) (tests.py:26)."""
    b0, b1 = basis(2, 0), basis(2, 1)
    s = 1 / np.sqrt(2)
    if state == "00":
        return s * (tensor(b0, b0) + tensor(b1, b1))
    if state == "01":
        return s * (tensor(b0, b0) - tensor(b1, b1))
    if state == "10":
        return s * (tensor(b0, b1) + tensor(b1, b0))
    if state == "11":
        return s * (tensor(b0, b1) - tensor(b1, b0))
    raise ValueError(state)


def rand_unitary_haar(N=2, dims=None, seed=None):
    rs = np.random if seed is None else np.random.RandomState(seed)
    z = (rs.normal(size=(N, N)) + 1j * rs.normal(size=(N, N))) / np.sqrt(2)
    q, r = np.linalg.qr(z)
    ph = np.diag(r) / np.abs(np.diag(r))
    u = q * ph
    return Qobj(u, dims=dims if dims is not None else [[N], [N]])


def rand_ket_haar(N=2, dims=None, seed=None):
    u = rand_unitary_haar(N, None, seed)
    psi = u * basis(N, 0)
    if dims is not None:
        psi.dims = dims
    return psi


# ----- qip.operations -------------------------------------------------------------

def gate_expand_1toN(U, N, target):
    if N < 1:
        raise ValueError("integer N must be larger or equal to 1")
    if target >= N:
        raise ValueError("target must be integer < integer N")
    return tensor([identity(2)] * target + [U] + [identity(2)] * (N - target - 1))


def gate_expand_2toN(U, N, control=None, target=None, targets=None):
    """4x4 ``U`` acts with its first tensor factor on ``control`` and its second on
    ``target``; identity elsewhere (gates.py:595-597)."""
    if targets is not None:
        control, target = targets
    if control is None or target is None:
        raise ValueError("Specify value of control and target")
    if N < 2:
        raise ValueError("integer N must be larger or equal to 2")
    if control >= N or target >= N:
        raise ValueError("control and not target must be integer < integer N")
    if control == target:
        raise ValueError("target and not control cannot be equal")
    u = U.full().reshape(2, 2, 2, 2)       # [c', t', c, t]
    dim = 2 ** N
    rows, cols, vals = [], [], []
    sc, st = N - 1 - control, N - 1 - target
    idx = np.arange(dim)
    cb = (idx >> sc) & 1
    tb = (idx >> st) & 1
    base = idx & ~((1 << sc) | (1 << st))
    for c2 in (0, 1):
        for t2 in (0, 1):
            v = u[c2, t2, cb, tb]
            r = base | (c2 << sc) | (t2 << st)
            nz = v != 0
            rows.append(r[nz]); cols.append(idx[nz]); vals.append(v[nz])
    m = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(dim, dim), dtype=np.complex128)
    return Qobj._new(m, [[2] * N, [2] * N])


def rx(phi, N=None, target=0):
    if N is not None:
        return gate_expand_1toN(rx(phi), N, target)
    return Qobj([[np.cos(phi / 2), -1j * np.sin(phi / 2)],
                 [-1j * np.sin(phi / 2), np.cos(phi / 2)]], dims=[[2], [2]])


def ry(phi, N=None, target=0):
    if N is not None:
        return gate_expand_1toN(ry(phi), N, target)
    return Qobj([[np.cos(phi / 2), -np.sin(phi / 2)],
                 [np.sin(phi / 2), np.cos(phi / 2)]], dims=[[2], [2]])


def rz(phi, N=None, target=0):
    if N is not None:
        return gate_expand_1toN(rz(phi), N, target)
    return Qobj([[np.exp(-1j * phi / 2), 0],
                 [0, np.exp(1j * phi / 2)]], dims=[[2], [2]])


def x_gate(N=None, target=0):
    if N is not None:
        return gate_expand_1toN(x_gate(), N, target)
    return sigmax()


def y_gate(N=None, target=0):
    if N is not None:
        return gate_expand_1toN(y_gate(), N, target)
    return sigmay()


def z_gate(N=None, target=0):
    if N is not None:
        return gate_expand_1toN(z_gate(), N, target)
    return sigmaz()


def phasegate(theta, N=None, target=0):
    if N is not None:
        return gate_expand_1toN(phasegate(theta), N, target)
    return Qobj([[1, 0], [0, np.exp(1.0j * theta)]], dims=[[2], [2]])


def s_gate(N=None, target=0):
    if N is not None:
        return gate_expand_1toN(s_gate(), N, target)
    return Qobj([[1, 0], [0, 1j]], dims=[[2], [2]])


def t_gate(N=None, target=0):
    if N is not None:
        return gate_expand_1toN(t_gate(), N, target)
    return Qobj([[1, 0], [0, np.exp(1j * np.pi / 4)]], dims=[[2], [2]])


def snot(N=None, target=0):
    if N is not None:
        return gate_expand_1toN(snot(), N, target)
    return Qobj(np.array([[1, 1], [1, -1]]) / np.sqrt(2.0), dims=[[2], [2]])


def cnot(N=None, control=0, target=1):
    if (control == 1 and target == 0) and N is None:
        N = 2
    if N is not None:
        return gate_expand_2toN(cnot(), N, control, target)
    return Qobj([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]],
                dims=[[2, 2], [2, 2]])


def csign(N=None, control=0, target=1):
    if (control == 1 and target == 0) and N is None:
        N = 2
    if N is not None:
        return gate_expand_2toN(csign(), N, control, target)
    return Qobj([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, -1]],
                dims=[[2, 2], [2, 2]])


cz_gate = csign


def sqrtiswap(N=None, targets=[0, 1]):
    if targets != [0, 1] and N is None:
        N = 2
    if N is not None:
        return gate_expand_2toN(sqrtiswap(), N, targets=targets)
    s = 1 / np.sqrt(2)
    return Qobj(np.array([[1, 0, 0, 0], [0, s, 1j * s, 0], [0, 1j * s, s, 0], [0, 0, 0, 1]]),
                dims=[[2, 2], [2, 2]])


# ----- module layout expected by `import qutip as qt` users ---------------------
def _submodule(name, **members):
    m = types.ModuleType(name)
    m.__dict__.update(members)
    return m


_ops = _submodule(
    __name__ + ".qip.operations",
    rx=rx, ry=ry, rz=rz, x_gate=x_gate, y_gate=y_gate, z_gate=z_gate,
    phasegate=phasegate, s_gate=s_gate, t_gate=t_gate, snot=snot, cnot=cnot,
    csign=csign, cz_gate=cz_gate, sqrtiswap=sqrtiswap,
    gate_expand_1toN=gate_expand_1toN, gate_expand_2toN=gate_expand_2toN)
qip = _submodule(__name__ + ".qip", operations=_ops)
states = _submodule(__name__ + ".states", bell_state=bell_state, basis=basis)
random_objects = _submodule(__name__ + ".random_objects",
                            rand_unitary_haar=rand_unitary_haar, rand_ket_haar=rand_ket_haar)


def install_as_qutip():
    """Register this module as ``qutip`` so the unmodified reference imports it.
    Only golden-generation / oracle self-tests call this."""
    me = sys.modules[__name__]
    sys.modules["qutip"] = me
    sys.modules["qutip.qip"] = qip
    sys.modules["qutip.qip.operations"] = _ops
    sys.modules["qutip.states"] = states
    sys.modules["qutip.random_objects"] = random_objects
    return me
