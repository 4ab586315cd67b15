"""Reference-side binding of libpqc_b200.so -- the file a maintainer of rmdocherty/pyramaterised
would add next to circuit.py to keep the reference's own classes (QuTiP gate objects, PQC) and
replace only the arithmetic of the hot path (level B).

It touches nothing but public attributes of the REFERENCE's objects (class names and the fields
set in pyramaterised/gates.py), the C ABI of include/pqc_b200.h, and torch for device buffers.
tests/test_reference_binding.py runs `lower` on circuits built with the unmodified reference and
checks that the op list equals the one pyramaterised_b200's own classes produce.
"""
import ctypes as C
import os

import numpy as np

LIB_PATH = os.environ.get("PQC_B200_LIB") or os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyramaterised_b200", "libpqc_b200.so")

# enum pqc_opcode (include/pqc_b200.h)
RX, RY, RZ, H, X, S, T, CNOT, CZ, SQRTISWAP, RXX, RYY, RZZ, FSIM, FIXED_FSIM, IDENT = range(16)


class PqcOp(C.Structure):                    # struct pqc_op
    _fields_ = [("kind", C.c_int32), ("q0", C.c_int32), ("q1", C.c_int32), ("param", C.c_int32),
                ("param2", C.c_int32), ("group", C.c_int32), ("scale", C.c_double),
                ("offset", C.c_double)]


_ONE_QUBIT = {"R_x": RX, "R_y": RY, "R_z": RZ, "I": IDENT, "negative_R_z": RZ, "offset_R_z": RZ}
_FIXED_1Q = {"fixed_R_y": RY, "fixed_R_z": RZ}
_CONST_1Q = {"H": H, "X": X, "S": S, "T": T}
_ENTANGLER = {"CNOT": CNOT, "CPHASE": CZ, "CZ": CZ, "sqrtiSWAP": SQRTISWAP}
_PAIR_ROT = {"R_xx": RXX, "R_yy": RYY, "R_zz": RZZ}


def _chain_pairs(N):                         # gates.py:366-371
    return [(2 * j, 2 * j + 1) for j in range(N // 2)] + \
           [(2 * j + 1, 2 * j + 2) for j in range((N - 1) // 2)]


def _all_pairs(N):                           # gates.py:393-394 (itertools.permutations order)
    return [(a, b) for a in range(N) for b in range(N) if a != b]


def _lower_gate(g, slot):
    """Primitive ops (kind, q0, q1, param, param2, scale, offset) of ONE reference gate object.
    `slot` = its first parameter slot, or -1 to bake its current angle in (param_count == 0)."""
    name = type(g).__name__
    if name in _ONE_QUBIT:                   # gates.py:106-205
        scale = -1.0 if name == "negative_R_z" else 1.0
        offset = float(getattr(g, "offset", 0.0)) if name == "offset_R_z" else 0.0
        if slot < 0:
            return [(_ONE_QUBIT[name], g.q_on, -1, -1, -1, 1.0, float(g.theta))]
        return [(_ONE_QUBIT[name], g.q_on, -1, slot, -1, scale, offset)]
    if name in _FIXED_1Q:                    # gates.py:252-283: angle frozen at construction
        return [(_FIXED_1Q[name], g.q_on, -1, -1, -1, 1.0, float(g.theta))]
    if name in _CONST_1Q:                    # gates.py:211-300
        return [(_CONST_1Q[name], g.q_on, -1, -1, -1, 1.0, 0.0)]
    if name in _ENTANGLER:                   # gates.py:327-349
        return [(_ENTANGLER[name], g.q1, g.q2, -1, -1, 1.0, 0.0)]
    if name in ("CHAIN", "ALLTOALL"):        # gates.py:355-401: first listed pair acts first
        kind = _ENTANGLER[g.entangler.__name__]
        pairs = _chain_pairs(g.q_N) if name == "CHAIN" else _all_pairs(g.q_N)
        return [(kind, a, b, -1, -1, 1.0, 0.0) for a, b in pairs]
    if name in _PAIR_ROT:                    # gates.py:492-551
        if slot < 0:
            return [(_PAIR_ROT[name], g.q1, g.q2, -1, -1, 1.0, float(g.theta))]
        return [(_PAIR_ROT[name], g.q1, g.q2, slot, -1, 1.0, 0.0)]
    if name in ("shared_parameter", "RR_block"):   # gates.py:441-484,554-585: one slot, every member
        ops = []
        for m in g.layer:
            ops += _lower_gate(m, slot)
        return ops
    if name == "fSim":                       # gates.py:651-698: two consecutive slots (theta, phi)
        if slot < 0:
            return [(FSIM, g.q1, g.q2, -1, -1, float(g.phi), float(g.theta))]
        return [(FSIM, g.q1, g.q2, slot, slot + 1, 1.0, 0.0)]
    if name == "fixed_fSim":                 # gates.py:740-759
        if slot < 0:
            return [(FIXED_FSIM, g.q1, g.q2, -1, -1, 1.0, float(g.theta))]
        return [(FIXED_FSIM, g.q1, g.q2, slot, -1, 1.0, 0.0)]
    raise NotImplementedError(f"no primitive lowering for reference gate {name}")


def lower(pqc):
    """PQC.gates (circuit.py:53-61) -> list of (kind, q0, q1, param, param2, group, scale,
    offset); group = index of the gate object in pqc.gates, parameter slots in gate order
    (circuit.py:86-116: a two-parameter gate takes two consecutive angles)."""
    ops, slot = [], 0
    for gi, g in enumerate(pqc.gates):
        for (kind, q0, q1, p, p2, scale, offset) in _lower_gate(g, slot if g.param_count > 0 else -1):
            ops.append((kind, q0, q1, p, p2, gi, scale, offset))
        slot += g.param_count
    return ops


# ---- calling the library ------------------------------------------------------------------------
_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.pqc_last_error.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc:
        raise RuntimeError(_load().pqc_last_error().decode())


def run_batch(pqc, angles):
    """Replaces the loop of PQC.run (circuit.py:118-125) for a batch angles[S, P]: returns a
    torch complex128 device tensor [S, 2^n] (wrap rows in qt.Qobj for Qobj semantics)."""
    import torch
    lib = _load()
    ops = lower(pqc)
    arr = (PqcOp * len(ops))(*[PqcOp(*o) for o in ops])
    a = torch.as_tensor(np.asarray(angles), dtype=torch.float64, device="cuda").contiguous()
    h = C.c_void_p()
    _check(lib.pqc_program_create(C.c_int(pqc.n_qubits), C.c_int(a.shape[1]), C.c_int(len(ops)), arr,
                                  C.byref(h)))
    try:
        init = torch.as_tensor(np.asarray(pqc.initial_state.full())[:, 0], dtype=torch.complex128,
                               device="cuda").contiguous()
        out = torch.empty((a.shape[0], 2 ** pqc.n_qubits), dtype=torch.complex128, device="cuda")
        _check(lib.pqc_run_batch(h, C.c_void_p(a.data_ptr()), C.c_int64(a.shape[1]),
                                 C.c_int64(a.shape[0]), C.c_void_p(init.data_ptr()), C.c_int64(0),
                                 C.c_void_p(out.data_ptr()),
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    finally:
        lib.pqc_program_destroy(h)
    return out


def meyer_wallach(states, n_qubits):
    """Replaces single_Q per state (measure.py:226-237) for a device batch [S, 2^n]."""
    import torch
    Q = torch.empty((states.shape[0],), dtype=torch.float64, device=states.device)
    _check(_load().pqc_meyer_wallach(C.c_void_p(states.data_ptr()), C.c_int64(states.shape[0]),
                                     C.c_int(n_qubits), C.c_void_p(Q.data_ptr()),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return Q
