"""ctypes binding of include/pqc_b200.h (libpqc_b200.so, built in-tree by build.py).

There is no CPU fallback: if the shared library is missing this module raises, and every
compute entry point needs a CUDA device.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PQC_LIB_PATH: developer override used by tools/microbench.py to A/B kernel build variants.
LIB_PATH = os.environ.get("PQC_LIB_PATH") or os.path.join(HERE, "libpqc_b200.so")

# opcodes (enum pqc_opcode)
OP_RX, OP_RY, OP_RZ, OP_H, OP_X, OP_S, OP_T, OP_CNOT, OP_CZ, OP_SQRTISWAP, OP_RXX, OP_RYY, \
    OP_RZZ, OP_FSIM, OP_FIXED_FSIM, OP_IDENT = range(16)
OP_NAMES = ["RX", "RY", "RZ", "H", "X", "S", "T", "CNOT", "CZ", "SQRTISWAP", "RXX", "RYY", "RZZ",
            "FSIM", "FIXED_FSIM", "IDENT"]


class PqcOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("q0", C.c_int32), ("q1", C.c_int32), ("param", C.c_int32),
                ("param2", C.c_int32), ("group", C.c_int32), ("scale", C.c_double),
                ("offset", C.c_double)]


class PqcPauliTerm(C.Structure):
    _fields_ = [("xmask", C.c_uint32), ("zmask", C.c_uint32), ("re", C.c_double),
                ("im", C.c_double)]


_P = C.c_void_p
_I64 = C.c_int64
_INT = C.c_int
_DBL = C.c_double

# name -> argtypes; every function returns int except pqc_last_error
SIGNATURES = {
    "pqc_abi_version": [],
    "pqc_profile_begin": [],
    "pqc_profile_end": [C.POINTER(_DBL)],
    "pqc_profile_kinds": [C.POINTER(_DBL), _INT],
    "pqc_device_check": [C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_INT)],
    "pqc_program_create": [_INT, _INT, _INT, C.POINTER(PqcOp), C.POINTER(_P)],
    "pqc_program_destroy": [_P],
    "pqc_program_stats": [_P, C.POINTER(_I64)],
    "pqc_program_describe": [_P, C.c_char_p, _I64],
    "pqc_run_batch": [_P, _P, _I64, _I64, _P, _I64, _P, _P],
    "pqc_gradients_batch": [_P, _P, _I64, _I64, _P, _I64, _P, _P],
    "pqc_qfim_from_grads": [_P, _P, _INT, _INT, _I64, _P, _P],
    "pqc_qfim_workspace_bytes": [_P, _I64, C.POINTER(_I64)],
    "pqc_qfim_batch": [_P, _P, _I64, _I64, _P, _P, _I64, _P, _P, _P],
    "pqc_eigvalsh_batch": [_P, _I64, _INT, _P, _P],
    "pqc_eigh_batch": [_P, _I64, _INT, _P, _P, _P],
    "pqc_count_greater": [_P, _I64, _INT, _DBL, _P, _P],
    "pqc_meyer_wallach": [_P, _I64, _INT, _P, _P],
    "pqc_ptrace_1q": [_P, _INT, _INT, _P, _P],
    "pqc_overlap_batch": [_P, _I64, _P, _I64, _I64, _I64, _P, _P],
    "pqc_fidelity_hist": [_P, _I64, _P, _I64, _INT, _INT, _I64, _P, _P, _P],
    "pqc_hist_f64": [_P, _I64, _I64, _P, _P],
    "pqc_kl_haar": [_P, _I64, _DBL, _P, _P, _P],
    "pqc_magic_batch": [_P, _I64, _INT, _INT, C.POINTER(_DBL), _P, _P],
    "pqc_pauli_expect_batch": [_P, _I64, _INT, _INT, C.POINTER(PqcPauliTerm), _P, _P],
    "pqc_pauli_apply_batch": [_P, _I64, _INT, _INT, C.POINTER(PqcPauliTerm), _P, _P],
    "pqc_dense_apply_batch": [_P, _I64, _INT, _P, _P, _P, _I64, _INT, _P, _P],
}

_lib = None


class PqcError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA engine with "
            "`python -m pyramaterised_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.pqc_last_error.restype = C.c_char_p
    lib.pqc_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = C.c_int
        fn.argtypes = args
    lib.pqc_launch_count.restype = C.c_longlong
    lib.pqc_launch_count.argtypes = []
    if lib.pqc_abi_version() != 1:
        raise ImportError("libpqc_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().pqc_last_error().decode()
        if msg == "No parameters supplied!":         # circuit.py:103,114
            raise Exception(msg)
        if msg.startswith("`bins` must be positive"):  # np.histogram (measure.py:153)
            raise ValueError(msg)
        raise PqcError(f"libpqc_b200 error {rc}: {msg}")
