"""Circuit factories with the reference's names and return values
(/root/reference/pyramaterised/templates.py).  Pure host-side list building: every
function returns a list of layers (lists of gate objects) for ``PQC.add_layer``.
"""
import random

import numpy as np

from .circuit import *          # noqa: F401,F403
from .circuit import PQC
from .gates import (CHAIN, CNOT, CPHASE, CZ, H, R_x, R_xx, R_y, R_yy, R_z, R_zz, RR_block, S, X,
                    EntGate, PRot, fSim, fixed_fSim, fixed_R_y, fixed_R_z, negative_R_z,
                    offset_R_z, shared_parameter, sqrtiSWAP)
from . import qobj as qt

LEN_CLIFF_STRING = 3


# ============================== HE circuits ==============================
def clifford_circuit_layers(p, N, method='random'):
    """Random Clifford layers (templates.py:17-48); uses Python's `random` like the reference."""
    pool = [H, S, CZ, CNOT]
    qubits = list(range(N))
    layers = []
    for _ in range(p):
        layer = []
        if method.lower() == 'random':
            for _n in range(N):
                gate = random.choice(pool)
                if issubclass(gate, PRot):
                    layer.append(gate(random.randint(0, N - 1), N))
                elif issubclass(gate, EntGate):
                    a, b = random.sample(qubits, k=2)
                    layer.append(gate([a, b], N))
        else:
            chosen = random.sample(qubits, k=N // 2)
            strings = [random.choices(pool[:2], k=LEN_CLIFF_STRING) for _q in chosen]
            for q, string in zip(chosen, strings):
                layer += [gate(q, N) for gate in string]
            layer.append(CHAIN(CNOT, N))
        layers.append(layer)
    return layers


def gen_shift_list(p, N):
    """a_s sequence of the NPQC paper: a_{s+q} = a_q (templates.py:51-63)."""
    pending = list(range(N // 2))
    shifts = np.zeros(2 ** (N // 2), dtype=np.int32)
    s = 1
    while pending:
        shifts[s - 1] = pending.pop(0)
        for q in range(1, s):
            shifts[s + q - 1] = shifts[q - 1]
        s *= 2
    return shifts


def NPQC_layers(p, N):
    """NPQC: identity QFIM at the returned reference angles (templates.py:66-95).
    Returns (layers, theta_ref)."""
    layers = [[R_y(i, N) for i in range(N)] + [R_z(i, N) for i in range(N)]]
    theta_ref = [np.pi / 2] * N + [0] * N
    shifts = gen_shift_list(p, N)
    evens = [2 * k - 2 for k in range(1, 1 + N // 2)]
    for i in range(p - 1):
        a_l = shifts[i]
        layer = [fixed_R_y(q, N, np.pi / 2) for q in evens]
        layer += [CPHASE([q, ((q + 1) + 2 * a_l) % N], N) for q in evens]
        for q in evens:
            layer += [R_y(q, N), R_z(q, N)]
            theta_ref += [np.pi / 2, 0]
        layers.append(layer)
    return layers, theta_ref


def string_to_entangler(string):
    table = {"cnot": CNOT, "cphase": CPHASE, "cz": CZ, "sqrtiswap": sqrtiSWAP}
    try:
        return table[string.lower()]
    except KeyError:
        raise Exception("Must supply a valid entangler!")


def _repeat(layer, p):
    return [layer for _ in range(p)]


def circuit_1_layers(p, N):
    """arXiv:1905.10876 circuit 1 (templates.py:114-119)."""
    return _repeat([R_x(i, N) for i in range(N)] + [R_y(i, N) for i in range(N)], p)


def circuit_2_layers(p, N, ent_str="cnot"):
    ent = string_to_entangler(ent_str)
    return _repeat([R_x(i, N) for i in range(N)] + [R_z(i, N) for i in range(N)] +
                   [CHAIN(ent, N)], p)


def circuit_9_layers(p, N, ent_str="cphase"):
    ent = string_to_entangler(ent_str)
    return _repeat([H(i, N) for i in range(N)] + [CHAIN(ent, N)] +
                   [R_x(i, N) for i in range(N)], p)


def qg_circuit_layers(p, N, ent_str="cnot"):
    """arXiv:2102.01659 (templates.py:140-151)."""
    ent = string_to_entangler(ent_str)
    block = []
    for rot in (R_z, R_x, R_z):
        block += [rot(i, N) for i in range(N)] + [CHAIN(ent, N)]
    return [[fixed_R_y(i, N, np.pi / 4) for i in range(N)]] + _repeat(block, p)


def generic_HE_layers(p, N, ent_str="cnot"):
    """Fixed ry(pi/4) layer, then p x [R_y, R_z, CHAIN] (templates.py:154-162)."""
    ent = string_to_entangler(ent_str)
    layer = [R_y(i, N) for i in range(N)] + [R_z(i, N) for i in range(N)] + [CHAIN(ent, N)]
    return [[fixed_R_y(i, N, np.pi / 4) for i in range(N)]] + _repeat(layer, p)


def clifford_HE_layers(p, N, ent_str="cnot"):
    ent = string_to_entangler(ent_str)
    return _repeat([R_y(i, N) for i in range(N)] + [R_z(i, N) for i in range(N)] +
                   [CHAIN(ent, N)], p)


def y_CPHASE_layers(p, N):
    return _repeat([R_y(i, N) for i in range(N)] + [CHAIN(CPHASE, N)], p)


def double_y_CPHASE_layers(p, N):
    return _repeat([R_y(i, N) for i in range(N)] + [R_y(i, N) for i in range(N)] +
                   [CHAIN(CPHASE, N)], p)


# ============================== problem-inspired circuits ==============================
def TFIM_layers(p, N):
    """H layer, then p x [ring of R_zz sharing one angle, R_x layer sharing one angle]
    (templates.py:194-204)."""
    layers = [[H(i, N) for i in range(N)]]
    for _ in range(p):
        layers.append([RR_block(R_zz, N), shared_parameter([R_x(i, N) for i in range(N)], N)])
    return layers


def modified_TFIM_layers(p, N):
    layers = []
    for _ in range(p):
        layers.append([RR_block(R_zz, N),
                       shared_parameter([R_x(i, N) for i in range(N)], N),
                       shared_parameter([R_z(i, N) for i in range(N)], N)])
    return layers


def TFIM_hamiltonian(N, g, h=0):
    """-sum_i (Z_i Z_{i+1} + g X_i + h Z_i), periodic (templates.py:220-227)."""
    Hm = 0
    for i in range(N):
        j = (i + 1) % N
        Hm += genFockOp(qt.sigmaz(), i, N) * genFockOp(qt.sigmaz(), j, N) \
            + g * genFockOp(qt.sigmax(), i, N) + h * genFockOp(qt.sigmaz(), i, N)
    return -1 * Hm


def XXZ_layers(p, N, commute=False):
    """XXZ ansatz (templates.py:230-257): per layer ZZ(odd), YY+XX(odd), ZZ(even), YY+XX(even)."""
    even = [(2 * i - 2, 2 * i - 1) for i in range(1, N // 2 + 1)]
    odd = [(2 * i - 1, (2 * i) % N) for i in range(1, N // 2 + 1)]

    def block(rots, pairs, **kw):
        return shared_parameter([r((a, b), N) for r in rots for a, b in pairs], N, **kw)

    layers = []
    for _ in range(p):
        layers.append([block([R_zz], odd), block([R_yy, R_xx], odd, commute=commute),
                       block([R_zz], even), block([R_yy, R_xx], even, commute=commute)])
    return layers


def gen_theta_block(q1, q2, N):
    """Fermionic 'theta block' (templates.py:260-274)."""
    return [sqrtiSWAP([q1, q2], N),
            shared_parameter([negative_R_z(q1, N), offset_R_z(q2, N, np.pi)], N),
            sqrtiSWAP([q1, q2], N),
            fixed_R_z(q2, N, np.pi)]


def list_to_pairs(x):
    return [(x[i], x[i + 1]) for i in range(0, len(x) - 1, 2)]


def fermionic_circuit_layers(p, N):
    """Diamond arrangement of theta blocks (templates.py:282-304)."""
    layers = []
    for _ in range(p):
        rows = []
        for m in range(1, 1 + N // 2):
            left = list(range(N // 2 - m + 1, N // 2 + 1))
            right = list(range(1 + N // 2, 1 + N // 2 + m))
            rows.append(list_to_pairs(left + right))
        for row in rows + rows[-2::-1]:
            layer = []
            for a, b in row:
                layer += gen_theta_block(a - 1, b - 1, N)
            layers.append(layer)
    return layers


def fSim_circuit_layers(p, N, rotator='y', fixed=False):
    """Rotations + fSim gates with periodic boundary (templates.py:307-353)."""
    rots = {'y': R_y, 'x': R_x, 'z': R_z}
    if rotator.lower() not in rots:
        raise Exception("Please supply a valid single qubit rotator")
    rot = rots[rotator.lower()]
    two = fixed_fSim if fixed else fSim
    layers = []
    for l in range(p):
        layer = [rot(i, N) for i in range(N)]
        if N % 2 == 0:
            layer += [two([i, (i + 1) % N], N) for i in range(l % 2, N, 2)]
        else:
            offset = l % N
            rest = [i for i in range(N) if i != offset]
            pairs = []
            if offset % 2 == 1:
                pairs.append((rest.pop(0), rest.pop(-1)))
            pairs += [(rest[i], rest[i + 1]) for i in range(0, len(rest), 2)]
            layer += [two([a, b], N) for a, b in pairs]
            layer.append(rot(offset, N))
        layers.append(layer)
    return layers


def add_layers(circuit, layers):
    for l in layers:
        circuit.add_layer(l)
    return circuit


def _half_filled_index(N, shuffle):
    """Basis index of |1>^{N/2} |0>^{N/2}, optionally shuffled with Python's `random`
    (templates.py:373-377)."""
    bits = [1] * (N // 2) + [0] * (N - N // 2)
    if shuffle:
        random.shuffle(bits)
    index = 0
    for b in bits:
        index = (index << 1) | b
    return index


def generate_circuit(circuit_type, N, p, hamiltonian="ZZ", rotator='', shuffle=True):
    """N qubit, p layer circuit from its name (templates.py:361-429)."""
    circuit = PQC(N)
    half_filled = False
    if circuit_type == "NPQC":
        layers, _theta_ref = NPQC_layers(p, N)
    elif circuit_type == "TFIM":
        layers = TFIM_layers(p, N)
    elif circuit_type == "TFIM_modified":
        layers = modified_TFIM_layers(p, N)
    elif circuit_type == "XXZ":
        layers, half_filled = XXZ_layers(p, N), True
    elif circuit_type == "Circuit_1":
        layers = circuit_1_layers(p, N)
    elif circuit_type == "Circuit_2":
        layers = circuit_2_layers(p, N)
    elif circuit_type == "Circuit_9":
        layers = circuit_9_layers(p, N)
    elif circuit_type == "qg_circuit":
        layers = qg_circuit_layers(p, N)
    elif circuit_type == "generic_HE":
        layers = generic_HE_layers(p, N)
    elif circuit_type == "clifford":
        layers = clifford_HE_layers(p, N)
    elif circuit_type == "y_CPHASE":
        layers = y_CPHASE_layers(p, N)
    elif circuit_type == "double_y_CPHASE":
        layers = double_y_CPHASE_layers(p, N)
    elif circuit_type == "fermionic":
        layers, half_filled = fermionic_circuit_layers(p, N), True
    elif circuit_type == "zfsim":
        layers, half_filled = fSim_circuit_layers(p, N, rotator='z'), True
    elif circuit_type == "fsim":
        layers = fSim_circuit_layers(p, N, rotator=rotator) if rotator in ['x', 'y', 'z'] \
            else fSim_circuit_layers(p, N)
        half_filled = True
    elif circuit_type == "fixed_fsim":
        layers, half_filled = fSim_circuit_layers(p, N, rotator='z', fixed=True), True
    else:
        raise UnboundLocalError("cannot access local variable 'layers'")   # as the reference fails
    if half_filled:
        circuit.set_initial_basis_state(_half_filled_index(N, shuffle))
    for l in layers:
        circuit.add_layer(l)
    return circuit
