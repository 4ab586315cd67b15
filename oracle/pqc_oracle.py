"""pqc_oracle -- CPU (numpy) restatement of the reference's PQC hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by pyramaterised_b200.

It restates, matrix-free, what rmdocherty/pyramaterised computes with 2^n x 2^n
QuTiP operators (citations are /root/reference paths):

  gate matrices and derivatives      pyramaterised/gates.py:106-205,211-300,327-349,
                                     355-401,441-585,588-759
  flat gate list, parameter order    pyramaterised/circuit.py:53-116
  |psi> = U_G ... U_1 |init>         pyramaterised/circuit.py:118-125
  derivative states                  pyramaterised/circuit.py:149-192
  QFIM, eigenvalues, EQD             pyramaterised/measure.py:33-99
  fidelities, histogram, KL          pyramaterised/measure.py:123-197
  Meyer-Wallach Q                    pyramaterised/measure.py:226-249
  Renyi / GKP magic                  pyramaterised/measure.py:268-368
  efficient_measurements             pyramaterised/measure.py:370-459
  circuit templates                  pyramaterised/templates.py:51-95,114-257,260-353,361-429

QuTiP conventions: qubit 0 is the most significant bit of
the basis index; for two-qubit matrices the first listed qubit is the more
significant of the 4x4 index.

Parity pinning: tests/test_oracle_golden.py checks every function here against
tests/golden/ref_golden.npz, which was produced by the unmodified reference
package (see tests/golden/make_golden.py), and against the reference's own
known answers (tests.py:64-87,114-128,192-212,284-295).

A circuit is a list of *gate specs* in the reference's vocabulary:

  ("R_x"|"R_y"|"R_z"|"negative_R_z"|"I", q)          one parameter
  ("offset_R_z", q, offset)                          one parameter
  ("H"|"X"|"S"|"T", q)                               fixed
  ("fixed_R_y"|"fixed_R_z", q, theta)                fixed
  ("CNOT"|"CPHASE"|"CZ"|"sqrtiSWAP", q1, q2)         fixed
  ("CHAIN"|"ALLTOALL", entangler_name)               fixed block
  ("R_zz"|"R_xx"|"R_yy", q1, q2)                     one parameter
  ("shared_parameter", [member specs], commute)      one parameter
  ("RR_block", rotator_name)                         one parameter
  ("fSim", q1, q2)                                   two parameters
  ("fixed_fSim", q1, q2)                             one parameter
"""
from __future__ import annotations

from itertools import permutations

import numpy as np
import scipy.linalg
import scipy.special

SX = np.array([[0, 1], [1, 0]], dtype=np.complex128)
SY = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
SZ = np.array([[1, 0], [0, -1]], dtype=np.complex128)
I2 = np.eye(2, dtype=np.complex128)
_PAULI = {"x": SX, "y": SY, "z": SZ}

_ROT1 = {"R_x": "x", "R_y": "y", "R_z": "z", "negative_R_z": "z", "offset_R_z": "z",
         "fixed_R_y": "y", "fixed_R_z": "z"}
_ROT2 = {"R_xx": "x", "R_yy": "y", "R_zz": "z"}
_ENT = ("CNOT", "CPHASE", "CZ", "sqrtiSWAP")


# ----------------------------------------------------------------------------------
# matrices (gates.py; QuTiP forms per SURVEY.md 8 a3-a8)
# ----------------------------------------------------------------------------------
def rot_matrix(axis, theta):
    """rx/ry/rz(theta) of qutip.qip.operations (gates.py:159-174)."""
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    if axis == "x":
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if axis == "y":
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    return np.array([[np.exp(-1j * theta / 2), 0], [0, np.exp(1j * theta / 2)]],
                    dtype=np.complex128)


def fixed1_matrix(name):
    if name == "H":      # x_gate * ry(pi/2)  (gates.py:226-232)
        return SX @ rot_matrix("y", np.pi / 2)
    if name == "X":      # gates.py:245-249
        return SX.copy()
    if name == "S":      # phasegate(pi/2)  (gates.py:286-291)
        return np.array([[1, 0], [0, np.exp(1j * np.pi / 2)]], dtype=np.complex128)
    if name == "T":      # gates.py:294-300
        return np.array([[1, 0], [0, np.exp(1j * np.pi / 4)]], dtype=np.complex128)
    raise KeyError(name)


def ent_matrix(name):
    if name == "CNOT":                  # gates.py:327-330
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]],
                        dtype=np.complex128)
    if name in ("CPHASE", "CZ"):        # gates.py:333-337,346-349 (quirk Q11)
        return np.diag([1, 1, 1, -1]).astype(np.complex128)
    if name == "sqrtiSWAP":             # gates.py:340-343
        s = 1 / np.sqrt(2)
        return np.array([[1, 0, 0, 0], [0, s, 1j * s, 0], [0, 1j * s, s, 0], [0, 0, 0, 1]],
                        dtype=np.complex128)
    raise KeyError(name)


def rr_matrix(axis, theta):
    """cos(theta/2) I - i sin(theta/2) sigma(x)sigma  (gates.py:509-517)."""
    p = _PAULI[axis]
    return np.cos(theta / 2) * np.eye(4) - 1j * np.sin(theta / 2) * np.kron(p, p)


def fsim_matrix(theta, phi, d=0):
    """fsim_gate / fsim_gate_d_theta / fsim_gate_d_phi (gates.py:588-648); d selects
    0 = gate, 1 = 'd/dtheta', 2 = 'd/dphi' exactly as written there (quirk Q3)."""
    c, s = np.cos(theta), np.sin(theta)
    e = np.exp(-1j * phi)
    if d == 0:
        mid, corner = [[c, -1j * s], [-1j * s, c]], e
    elif d == 1:
        mid, corner = [[-s, -1j * c], [-1j * c, -s]], e
    else:
        mid, corner = [[c, -1j * s], [-1j * s, c]], -1j * e
    m = np.zeros((4, 4), dtype=np.complex128)
    m[0, 0] = 1
    m[1:3, 1:3] = mid
    m[3, 3] = corner
    return m


def fixed_fsim_matrix(theta, d=0):
    """fixed_fsim_gate / fixed_fsim_gate_d_theta (gates.py:700-737)."""
    c, s = np.cos(theta), np.sin(theta)
    mid = [[c, -1j * s], [-1j * s, c]] if d == 0 else [[-s, -1j * c], [-1j * c, -s]]
    m = np.zeros((4, 4), dtype=np.complex128)
    m[0, 0] = 1
    m[1:3, 1:3] = mid
    m[3, 3] = 1
    return m


# ----------------------------------------------------------------------------------
# matrix-free application on a batch psi[S, 2^n]  (a1/a2: gates.py:39-46,63-85)
# ----------------------------------------------------------------------------------
def apply_1q(psi, n, q, U):
    """U (2x2, or per-sample [S,2,2]) on qubit q; qubit 0 is the most significant bit."""
    S = psi.shape[0]
    v = psi.reshape(S, 2 ** q, 2, 2 ** (n - q - 1))
    if U.ndim == 2:
        out = np.einsum("ab,sxby->sxay", U, v)
    else:
        out = np.einsum("sab,sxby->sxay", U, v)
    return out.reshape(S, -1)


def apply_2q(psi, n, q1, q2, U):
    """4x4 U (or [S,4,4]) whose first tensor factor acts on q1, second on q2."""
    S = psi.shape[0]
    U = np.asarray(U)
    batched = U.ndim == 3
    u = U.reshape((S,) * batched + (2, 2, 2, 2))          # [.., a', b', a, b]
    if q1 > q2:                                           # make the first axis the lower qubit
        u = np.swapaxes(np.swapaxes(u, -4, -3), -2, -1)
        q1, q2 = q2, q1
    v = psi.reshape(S, 2 ** q1, 2, 2 ** (q2 - q1 - 1), 2, 2 ** (n - q2 - 1))
    if batched:
        out = np.einsum("sijkl,sxkylz->sxiyjz", u, v)
    else:
        out = np.einsum("ijkl,sxkylz->sxiyjz", u, v)
    return out.reshape(S, -1)


def chain_pairs(n):
    """CHAIN order of application: (0,1),(2,3),... then (1,2),(3,4),... (gates.py:366-376)."""
    return [(2 * j, 2 * j + 1) for j in range(n // 2)] + \
           [(2 * j + 1, 2 * j + 2) for j in range((n - 1) // 2)]


def alltoall_pairs(n):
    """ALLTOALL: every ordered pair (gates.py:392-398, quirk Q10)."""
    return list(permutations(range(n), 2))


def ring_pairs(n):
    """RR_block ring (i,(i+1) mod n) (gates.py:565-572, quirk Q9 for n=2)."""
    return [(i, (i + 1) % n) for i in range(n)]


# ----------------------------------------------------------------------------------
# gate spec semantics
# ----------------------------------------------------------------------------------
def param_count(spec):
    k = spec[0]
    if k in ("R_x", "R_y", "R_z", "negative_R_z", "offset_R_z", "I", "R_zz", "R_xx", "R_yy",
             "shared_parameter", "RR_block", "fixed_fSim"):
        return 1
    if k == "fSim":
        return 2
    return 0


def members(spec, n):
    """Member rotations of a shared-parameter block, in application order."""
    if spec[0] == "shared_parameter":
        return list(spec[1])
    if spec[0] == "RR_block":
        return [(spec[1], a, b) for a, b in ring_pairs(n)]
    raise KeyError(spec[0])


def _effective_angle(spec, theta):
    """set_theta semantics: negative_R_z stores -theta (gates.py:181-183), offset_R_z
    stores theta + offset (gates.py:203-205)."""
    if spec[0] == "negative_R_z":
        return -theta
    if spec[0] == "offset_R_z":
        return theta + spec[2]
    return theta


def _per_sample(fn, theta):
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    return np.stack([fn(t) for t in theta])


def apply_gate(psi, n, spec, theta=None, phi=None, conj=False):
    """g * psi (gates.py:63-67).  theta/phi are per-sample arrays [S] for parameterised
    gates.  conj=True applies the element-wise conjugated operator (needed for the
    shared_parameter(commute=False) derivative, gates.py:464-466)."""
    k = spec[0]
    cj = (lambda m: m.conj()) if conj else (lambda m: m)
    if k in ("R_x", "R_y", "R_z", "negative_R_z", "offset_R_z"):
        ax = _ROT1[k]
        U = _per_sample(lambda t: rot_matrix(ax, _effective_angle(spec, t)), theta)
        return apply_1q(psi, n, spec[1], cj(U))
    if k == "I":                      # gates.py:150-156 (quirk Q5): identity, eats a parameter
        return psi
    if k in ("fixed_R_y", "fixed_R_z"):
        return apply_1q(psi, n, spec[1], cj(rot_matrix(_ROT1[k], spec[2])))
    if k in ("H", "X", "S", "T"):
        return apply_1q(psi, n, spec[1], cj(fixed1_matrix(k)))
    if k in _ENT:
        return apply_2q(psi, n, spec[1], spec[2], cj(ent_matrix(k)))
    if k == "CHAIN":
        for a, b in chain_pairs(n):
            psi = apply_2q(psi, n, a, b, cj(ent_matrix(spec[1])))
        return psi
    if k == "ALLTOALL":
        for a, b in alltoall_pairs(n):
            psi = apply_2q(psi, n, a, b, cj(ent_matrix(spec[1])))
        return psi
    if k in _ROT2:
        ax = _ROT2[k]
        U = _per_sample(lambda t: rr_matrix(ax, t), theta)
        return apply_2q(psi, n, spec[1], spec[2], cj(U))
    if k in ("shared_parameter", "RR_block"):
        # operation = prod(layer[::-1]): first member acts first (gates.py:475-477,580-582)
        for m in members(spec, n):
            psi = apply_gate(psi, n, m, theta, conj=conj)
        return psi
    if k == "fSim":
        th = np.atleast_1d(theta)
        ph = np.atleast_1d(phi)
        U = np.stack([fsim_matrix(t, p) for t, p in zip(th, ph)])
        return apply_2q(psi, n, spec[1], spec[2], cj(U))
    if k == "fixed_fSim":
        U = _per_sample(lambda t: fixed_fsim_matrix(t), theta)
        return apply_2q(psi, n, spec[1], spec[2], cj(U))
    raise KeyError(f"unknown gate spec {spec!r}")


def apply_generator(psi, n, spec):
    """member.derivative() applied to psi: -i/2 * Pauli (gates.py:133-138,519-522);
    +i/2 for negative_R_z (gates.py:185-187)."""
    k = spec[0]
    if k in ("R_x", "R_y", "R_z", "offset_R_z"):
        return apply_1q(psi, n, spec[1], -0.5j * _PAULI[_ROT1[k]])
    if k == "negative_R_z":
        return apply_1q(psi, n, spec[1], 0.5j * SZ)
    if k == "I":
        return -0.5j * psi
    if k in _ROT2:
        p = _PAULI[_ROT2[k]]
        return apply_2q(psi, n, spec[1], spec[2], -0.5j * np.kron(p, p))
    raise KeyError(f"no generator for {spec!r}")


def apply_deriv_gate(psi, n, spec, theta=None, phi=None, which=0):
    """(deriv * gate) psi as take_derivative builds it (circuit.py:160-165)."""
    k = spec[0]
    if param_count(spec) == 0:
        # Gate.derivative() is the identity (gates.py:93-97); reachable only through Q2
        return apply_gate(psi, n, spec, theta, phi)
    if k == "fSim":
        d = 1 if which in (0, 1) else 2
        if which == 0:
            raise AttributeError("fSim has no derivative() for param=0")
        out = apply_gate(psi, n, spec, theta, phi)
        U = np.stack([fsim_matrix(t, p, d) for t, p in zip(np.atleast_1d(theta),
                                                           np.atleast_1d(phi))])
        return apply_2q(out, n, spec[1], spec[2], U)
    if k == "fixed_fSim":
        out = apply_gate(psi, n, spec, theta)
        U = _per_sample(lambda t: fixed_fsim_matrix(t, 1), theta)
        return apply_2q(out, n, spec[1], spec[2], U)
    if k in ("shared_parameter", "RR_block"):
        mem = members(spec, n)
        commute = True if k == "RR_block" else bool(spec[2])
        out = apply_gate(psi, n, spec, theta)
        if commute:                                   # gates.py:454-457
            acc = 0
            for m in mem:
                acc = acc + apply_generator(out, n, m)
            return acc
        # gates.py:458-466: sum_k prod(layer with member k -> D_k*U_k) . conj(op) . op
        out = apply_gate(out, n, spec, theta, conj=True)
        acc = 0
        for kk in range(len(mem)):
            v = out
            for j, m in enumerate(mem):
                v = apply_gate(v, n, m, theta)
                if j == kk:
                    v = apply_generator(v, n, m)
            acc = acc + v
        return acc
    out = apply_gate(psi, n, spec, theta, phi)
    return apply_generator(out, n, spec)


# ----------------------------------------------------------------------------------
# circuit level  (circuit.py)
# ----------------------------------------------------------------------------------
def n_params(specs):
    return sum(param_count(s) for s in specs)


def parameterised_attr(specs):
    """PQC.parameterised / PQC.n_params exactly as set_gates builds them, including the
    for...else that always runs (quirk Q1, circuit.py:62-72)."""
    out, total = [], 0
    for s in specs:
        c = param_count(s)
        total += c
        for _ in range(c):
            total += 1
            out.append(total)
        out.append(-1)
    return out, total


def zero_state(n):
    v = np.zeros(2 ** n, dtype=np.complex128)
    v[0] = 1
    return v


def _slots(specs):
    """first parameter slot of every gate, in gate order (circuit.py:89-116)."""
    slots, k = [], 0
    for s in specs:
        slots.append(k)
        k += param_count(s)
    return slots


def run(specs, n, angles, init=None):
    """PQC.run for a batch: angles[S,P] -> states[S,2^n] (circuit.py:118-125)."""
    angles = np.atleast_2d(np.asarray(angles, dtype=np.float64))
    S = angles.shape[0]
    if angles.shape[1] < n_params(specs):
        raise IndexError("list index out of range")       # circuit.py:97,110
    init = zero_state(n) if init is None else np.asarray(init, dtype=np.complex128)
    psi = np.tile(init, (S, 1))
    for s, k in zip(specs, _slots(specs)):
        c = param_count(s)
        th = angles[:, k] if c >= 1 else None
        ph = angles[:, k + 1] if c == 2 else None
        psi = apply_gate(psi, n, s, th, ph)
    return psi


def gradients(specs, n, angles, init=None, only=None):
    """PQC.get_gradients for a batch: [S,P,2^n] (circuit.py:149-192), including the
    quirk-Q2 gate lookup for two-parameter gates.  `only`: optional iterable of parameter
    indices to compute (each is an independent full re-simulation, circuit.py:167-169);
    the result then has len(only) derivative states."""
    angles = np.atleast_2d(np.asarray(angles, dtype=np.float64))
    S = angles.shape[0]
    init = zero_state(n) if init is None else np.asarray(init, dtype=np.complex128)
    slots = _slots(specs)

    def resim(loc, which):
        psi = np.tile(init, (S, 1))
        for j, (s, k) in enumerate(zip(specs, slots)):
            c = param_count(s)
            th = angles[:, k] if c >= 1 else None
            ph = angles[:, k + 1] if c == 2 else None
            if j == loc:
                psi = apply_deriv_gate(psi, n, s, th, ph, which)
            else:
                psi = apply_gate(psi, n, s, th, ph)
        return psi

    jobs = []
    param_locs = [j for j, s in enumerate(specs) if param_count(s) > 0]
    for count, loc in enumerate(param_locs):
        if param_count(specs[loc]) == 1:
            jobs.append((loc, 0))
        else:
            jobs.append((loc, 1))
            # circuit.py:188: g_prime = self.gates[count] -- count indexes the parameterised
            # list, so this is the intended gate only if every earlier gate is parameterised
            loc2 = count
            jobs.append((loc2, 2 if specs[loc2][0] == "fSim" else 0))
    if only is not None:
        jobs = [jobs[i] for i in only]
    out = [resim(loc, which) for loc, which in jobs]
    return np.stack(out, axis=1) if out else np.zeros((S, 0, 2 ** n), np.complex128)


def cost_zz(states):
    """<psi|Z0 Z1|psi>, the default Hamiltonian (circuit.py:28-31,132-137)."""
    states = np.atleast_2d(states)
    D = states.shape[1]
    n = D.bit_length() - 1
    idx = np.arange(D)
    sign = 1 - 2 * (((idx >> (n - 1)) ^ (idx >> (n - 2))) & 1)
    return (np.abs(states) ** 2 * sign).sum(axis=1)


# ----------------------------------------------------------------------------------
# measures  (measure.py)
# ----------------------------------------------------------------------------------
def qfi(state, grads):
    """get_QFI (measure.py:33-71): 4 Re(<d_p|d_q> - conj<psi|d_p> <psi|d_q>)."""
    grads = np.asarray(grads)
    s = grads @ state.conj()                    # <psi|d_p>
    G = grads.conj() @ grads.T                  # <d_p|d_q>
    F = 4 * np.real(G - np.outer(s.conj(), s))
    iu = np.triu_indices(len(F), 1)
    F[(iu[1], iu[0])] = F[iu]                   # mirror the upper triangle (measure.py:66-70)
    return F


def eqd(F, cutoff):
    """get_effective_quantum_dimension (measure.py:77-87)."""
    w = scipy.linalg.eigh(F, eigvals_only=True)
    return int(np.sum(w > cutoff))


def new_measure(F):
    """measure.py:89-99."""
    w = scipy.linalg.eigh(F, eigvals_only=True)
    return float(sum(1 if v > 1 else v for v in w))


def fidelity_samples(states):
    """_gen_f_samples (measure.py:123-137): |<psi_i|psi_j>|^2 for i<j in
    itertools.combinations order."""
    A = np.asarray(states)
    G = A.conj() @ A.T
    iu = np.triu_indices(A.shape[0], 1)
    return np.abs(G[iu]) ** 2


def gen_histo(F):
    """_gen_histo (measure.py:139-159); `filt` is a no-op there because the unfiltered
    list is what gets histogrammed (measure.py:150-154)."""
    F = np.asarray(F, dtype=np.float64)
    bins = int((75 / 10000) * len(F))
    counts, edges = np.histogram(F, bins=bins, range=(0, 1))
    prob = counts / counts.sum()
    mid = np.array([(edges[i - 1] + edges[i]) / 2 for i in range(1, len(edges))])
    return prob, mid, counts


def expr(F, N):
    """Measurements.expr (measure.py:161-180): KL(P_pqc || P_haar(N))."""
    if len(F) == 0:
        return 0
    prob, mid, _ = gen_histo(F)
    haar = (N - 1) * ((1 - mid) ** (N - 2))
    p_haar = haar / haar.sum()
    return float(np.sum(scipy.special.kl_div(prob, p_haar)))


def expr_from_counts(counts, N):
    """expr() restated on histogram counts (what the GPU path reduces across ranks)."""
    counts = np.asarray(counts, dtype=np.float64)
    bins = len(counts)
    edges = np.linspace(0, 1, bins + 1)
    mid = np.array([(edges[i - 1] + edges[i]) / 2 for i in range(1, len(edges))])
    prob = counts / counts.sum()
    haar = (N - 1) * ((1 - mid) ** (N - 2))
    p_haar = haar / haar.sum()
    return float(np.sum(scipy.special.kl_div(prob, p_haar)))


def single_Q(state, n):
    """Meyer-Wallach Q (measure.py:226-237)."""
    psi = np.asarray(state)
    tot = 0.0
    for k in range(n):
        m = psi.reshape(2 ** k, 2, 2 ** (n - k - 1)).transpose(1, 0, 2).reshape(2, -1)
        rho = m @ m.conj().T
        tot += np.real(np.trace(rho @ rho))
    return 2 * (1 - tot / n)


def conversion_matrices(n):
    """get_conversion_matrices (measure.py:268-313): xor[j,k] = j^k,
    sign[i,j] = (-1)^popcount(i&j)."""
    idx = np.arange(2 ** n)
    xor = idx[:, None] ^ idx[None, :]
    a = idx[:, None] & idx[None, :]
    par = np.zeros_like(a)
    for b in range(n):
        par ^= (a >> b) & 1
    return xor, 1 - 2 * par


def renyi_dense(state, alpha=2.0, conv=None):
    """renyi_entropy_fast, literal dense form (measure.py:335-349)."""
    c = np.asarray(state)
    n = c.shape[0].bit_length() - 1
    xor, sign = conversion_matrices(n) if conv is None else conv
    M = np.dot(np.conjugate(c) * sign, c[xor])
    r = np.sum(np.abs(2 ** (-n / 2) * M) ** (2 * alpha))
    return 1 / (1 - alpha) * np.log(r) - np.log(2 ** n)


def fwht(a):
    """Unnormalised Walsh-Hadamard transform along the last axis."""
    a = np.array(a, copy=True)
    D = a.shape[-1]
    h = 1
    while h < D:
        v = a.reshape(a.shape[:-1] + (D // (2 * h), 2, h))
        x, y = v[..., 0, :].copy(), v[..., 1, :].copy()
        v[..., 0, :] = x + y
        v[..., 1, :] = x - y
        h *= 2
    return a


def renyi_fwht(state, alpha=2.0):
    """Same quantity as renyi_dense: for every X-mask k the column M[:,k] is the
    Walsh-Hadamard transform over j of conj(c_j) c_{j^k}."""
    c = np.asarray(state)
    D = c.shape[0]
    n = D.bit_length() - 1
    idx = np.arange(D)
    tot = 0.0
    blk = max(1, min(D, (1 << 22) // D))
    for k0 in range(0, D, blk):
        ks = np.arange(k0, min(D, k0 + blk))
        v = np.conjugate(c)[None, :] * c[idx[None, :] ^ ks[:, None]]
        W = fwht(v)
        tot += np.sum(np.abs(2 ** (-n / 2) * W) ** (2 * alpha))
    return 1 / (1 - alpha) * np.log(tot) - np.log(2 ** n)


def renyi(state, alpha=2.0):
    D = len(state)
    return renyi_dense(state, alpha) if D <= 256 else renyi_fwht(state, alpha)


def gkp(state):
    """gkp_fast (measure.py:361-368)."""
    return 1 / (2 * np.log(2)) * renyi(state, alpha=0.5)


def efficient_measurements(states, n, measure_expr=True, measure_ent=True, measure_eom=True,
                           measure_GKP=True, full_data=False):
    """efficient_measurements on an already generated sample set (measure.py:385-459)."""
    S = len(states)
    if S == 0:
        measure_expr = measure_ent = measure_eom = measure_GKP = False
    overlaps, q_vals, magics, gkps = [], [], [], []
    if measure_expr and n < 12:
        overlaps = list(fidelity_samples(states)) if S > 1 else []
        e = expr(overlaps, 2 ** n) if n < 7 else -1
    else:
        e = -1
    if measure_ent:
        q_vals = [single_Q(s, n) for s in states]
        q, std = np.mean(q_vals), np.std(q_vals)
    else:
        q, std = -1, -1
    if measure_eom:
        magics = [renyi(s) for s in states]
        mb, ms = np.mean(magics), np.std(magics)
    else:
        mb, ms = -1, -1
    if measure_GKP:
        gkps = [gkp(s) for s in states]
        gb, gs = np.mean(gkps), np.std(gkps)
    else:
        gb, gs = -1, -1
    if full_data:
        return {"Expr": overlaps, "Ent": q_vals, "Magic": magics, "GKP": gkps}
    return {"Expr": e, "Ent": [q, std], "Magic": [mb, ms], "GKP": [gb, gs]}


# ----------------------------------------------------------------------------------
# templates  (templates.py) -> (specs, theta_ref or None, init)
# ----------------------------------------------------------------------------------
def gen_shift_list(N):
    """templates.py:51-63."""
    A = list(range(N // 2))
    s = 1
    shift = np.zeros(2 ** (N // 2), dtype=np.int64)
    while A:
        shift[s - 1] = A.pop(0)
        for q in range(1, s):
            shift[s + q - 1] = shift[q - 1]
        s *= 2
    return shift


def npqc(p, N):
    """NPQC_layers (templates.py:66-95) -> (specs, theta_ref)."""
    specs = [("R_y", i) for i in range(N)] + [("R_z", i) for i in range(N)]
    ref = [np.pi / 2] * N + [0.0] * N
    shift = gen_shift_list(N)
    for i in range(p - 1):
        a = int(shift[i])
        evens = [2 * k - 2 for k in range(1, 1 + N // 2)]
        specs += [("fixed_R_y", q, np.pi / 2) for q in evens]
        specs += [("CPHASE", q, ((q + 1) + 2 * a) % N) for q in evens]
        for q in evens:
            specs += [("R_y", q), ("R_z", q)]
            ref += [np.pi / 2, 0.0]
    return specs, ref


def _xxz_indices(N):
    even, odd = [], []
    for i in range(1, N // 2 + 1):
        even.append((2 * i - 2, 2 * i - 1))
        odd.append((2 * i - 1, (2 * i) % N))
    return even, odd


def _theta_block(q1, q2):
    """gen_theta_block (templates.py:260-274)."""
    return [("sqrtiSWAP", q1, q2),
            ("shared_parameter", [("negative_R_z", q1), ("offset_R_z", q2, np.pi)], True),
            ("sqrtiSWAP", q1, q2),
            ("fixed_R_z", q2, np.pi)]


def _fermionic(p, N):
    """fermionic_circuit_layers (templates.py:282-304)."""
    specs = []
    for _ in range(p):
        blocks = []
        for m in range(1, 1 + N // 2):
            first = list(range(N // 2, N // 2 - m, -1))[::-1]
            second = list(range(1 + N // 2, 1 + N // 2 + m))
            comb = first + second
            blocks.append([(comb[i], comb[i + 1]) for i in range(0, len(comb) - 1, 2)])
        blocks = blocks + list(reversed(blocks))[1:]
        for b in blocks:
            for x in b:
                specs += _theta_block(x[0] - 1, x[1] - 1)
    return specs


def _fsim_layers(p, N, rotator="y", fixed=False):
    """fSim_circuit_layers (templates.py:307-353)."""
    rot = {"x": "R_x", "y": "R_y", "z": "R_z"}[rotator.lower()]
    g = "fixed_fSim" if fixed else "fSim"
    specs = []
    for l in range(p):
        specs += [(rot, i) for i in range(N)]
        if N % 2 == 0:
            for i in range(l % 2, N, 2):
                specs.append((g, i, (i + 1) % N))
        else:
            offset = l % N
            idx = list(range(N))
            idx.pop(offset)
            pairs = []
            if offset % 2 == 1:
                bottom = idx.pop(0)
                top = idx.pop(-1)
                pairs.append((bottom, top))
            pairs += [(idx[i], idx[i + 1]) for i in range(0, len(idx), 2)]
            specs += [(g, a, b) for a, b in pairs]
            specs.append((rot, offset))
    return specs


def half_filled_state(N):
    """|1^{N/2} 0^{N/2}> -- generate_circuit(..., shuffle=False) (templates.py:373-377)."""
    idx = 0
    for q in range(N // 2):
        idx |= 1 << (N - 1 - q)
    v = np.zeros(2 ** N, dtype=np.complex128)
    v[idx] = 1
    return v


def generate_circuit(kind, N, p, rotator=""):
    """generate_circuit with shuffle=False (templates.py:361-429) -> (specs, init)."""
    init = None
    if kind == "NPQC":
        specs, _ = npqc(p, N)
    elif kind == "TFIM":                                   # templates.py:194-204
        specs = [("H", i) for i in range(N)]
        for _ in range(p):
            specs += [("RR_block", "R_zz"),
                      ("shared_parameter", [("R_x", i) for i in range(N)], True)]
    elif kind == "TFIM_modified":                          # templates.py:207-217
        specs = []
        for _ in range(p):
            specs += [("RR_block", "R_zz"),
                      ("shared_parameter", [("R_x", i) for i in range(N)], True),
                      ("shared_parameter", [("R_z", i) for i in range(N)], True)]
    elif kind == "XXZ":                                    # templates.py:230-257
        even, odd = _xxz_indices(N)
        specs = []
        for _ in range(p):
            specs += [
                ("shared_parameter", [("R_zz", a, b) for a, b in odd], True),
                ("shared_parameter", [("R_yy", a, b) for a, b in odd] +
                 [("R_xx", a, b) for a, b in odd], False),
                ("shared_parameter", [("R_zz", a, b) for a, b in even], True),
                ("shared_parameter", [("R_yy", a, b) for a, b in even] +
                 [("R_xx", a, b) for a, b in even], False)]
        init = half_filled_state(N)
    elif kind == "Circuit_1":
        specs = ([("R_x", i) for i in range(N)] + [("R_y", i) for i in range(N)]) * p
    elif kind == "Circuit_2":
        specs = ([("R_x", i) for i in range(N)] + [("R_z", i) for i in range(N)] +
                 [("CHAIN", "CNOT")]) * p
    elif kind == "Circuit_9":
        specs = ([("H", i) for i in range(N)] + [("CHAIN", "CPHASE")] +
                 [("R_x", i) for i in range(N)]) * p
    elif kind == "qg_circuit":                             # templates.py:140-151
        specs = [("fixed_R_y", i, np.pi / 4) for i in range(N)]
        for _ in range(p):
            for ax in ("R_z", "R_x", "R_z"):
                specs += [(ax, i) for i in range(N)] + [("CHAIN", "CNOT")]
    elif kind == "generic_HE":                             # templates.py:154-162
        specs = [("fixed_R_y", i, np.pi / 4) for i in range(N)]
        specs += ([("R_y", i) for i in range(N)] + [("R_z", i) for i in range(N)] +
                  [("CHAIN", "CNOT")]) * p
    elif kind == "clifford":                               # templates.py:164-171
        specs = ([("R_y", i) for i in range(N)] + [("R_z", i) for i in range(N)] +
                 [("CHAIN", "CNOT")]) * p
    elif kind == "y_CPHASE":
        specs = ([("R_y", i) for i in range(N)] + [("CHAIN", "CPHASE")]) * p
    elif kind == "double_y_CPHASE":
        specs = ([("R_y", i) for i in range(N)] * 2 + [("CHAIN", "CPHASE")]) * p
    elif kind == "fermionic":
        specs = _fermionic(p, N)
        init = half_filled_state(N)
    elif kind == "zfsim":
        specs = _fsim_layers(p, N, "z")
        init = half_filled_state(N)
    elif kind == "fsim":
        specs = _fsim_layers(p, N, rotator if rotator in ("x", "y", "z") else "y")
        init = half_filled_state(N)
    elif kind == "fixed_fsim":
        specs = _fsim_layers(p, N, "z", fixed=True)
        init = half_filled_state(N)
    else:
        raise KeyError(kind)
    return specs, init


def n_template_layers(kind, p):
    """Number of template layers incl. an initial fixed layer (the L in the
    algorithmic byte count L * 2 * 16 * 2^n)."""
    return p + 1 if kind in ("TFIM", "generic_HE", "qg_circuit") else p
