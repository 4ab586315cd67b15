"""Test helpers: oracle gate specs for the golden cases, and a converter from gate
*objects* (reference API attributes only: class name, q_on, q1/q2, theta, offset, layer,
commute, entangler, rotator -- gates.py) to oracle specs, so the oracle never sees the
product's own lowering."""
import numpy as np

from oracle import pqc_oracle as orc

_TEMPLATE = {
    "npqc_4_4": ("NPQC", 4, 4), "npqc_6_5": ("NPQC", 6, 5), "he_5_3": ("generic_HE", 5, 3),
    "he_10_10": ("generic_HE", 10, 10), "clifford_4_2": ("clifford", 4, 2),
    "tfim_4_4": ("TFIM", 4, 4), "tfim_6_3": ("TFIM", 6, 3),
    "tfimmod_4_2": ("TFIM_modified", 4, 2), "xxz_4_2": ("XXZ", 4, 2), "xxz_6_2": ("XXZ", 6, 2),
    "circuit1_4_2": ("Circuit_1", 4, 2), "circuit2_5_2": ("Circuit_2", 5, 2),
    "circuit9_4_2": ("Circuit_9", 4, 2), "ycphase_5_2": ("y_CPHASE", 5, 2),
    "dycphase_4_2": ("double_y_CPHASE", 4, 2), "qgt_4_2": ("qg_circuit", 4, 2),
    "fermionic_4_1": ("fermionic", 4, 1), "fermionic_6_1": ("fermionic", 6, 1),
    "zfsim_4_2": ("zfsim", 4, 2), "fsim_5_2": ("fsim", 5, 2, "x"),
    "fixedfsim_4_2": ("fixed_fsim", 4, 2),
}


def oracle_case(name):
    """-> (specs, n, init) for a golden case name (tests/golden/cases.py)."""
    if name in _TEMPLATE:
        t = _TEMPLATE[name]
        specs, init = orc.generate_circuit(t[0], t[1], t[2], *(t[3:]))
        return specs, t[1], init
    if name == "qg4":
        specs = [("fixed_R_y", i, np.pi / 4) for i in range(4)]
        specs += [("R_z", 0), ("R_x", 1), ("R_y", 2), ("R_z", 3), ("CHAIN", "CNOT")]
        specs += [("R_x", 0), ("R_x", 1), ("R_x", 2), ("R_y", 3), ("CHAIN", "CNOT")]
        specs += [("R_z", 0), ("R_x", 1), ("R_y", 2), ("R_y", 3), ("CHAIN", "CNOT")]
        return specs, 4, None
    if name == "example4":
        return ([("R_x", i) for i in range(4)] + [("CHAIN", "CNOT")]) * 3, 4, None
    if name == "mixed5":
        N = 5
        specs = [("H", 0), ("X", 1), ("S", 2), ("T", 3), ("fixed_R_z", 4, 0.3),
                 ("fixed_R_y", 2, 1.1)]
        specs += [("R_y", i) for i in range(N)] + [("CZ", 0, 3), ("CPHASE", 4, 1),
                                                   ("CNOT", 3, 1), ("CNOT", 0, 4)]
        specs += [("R_xx", 0, 2), ("R_yy", 3, 1), ("R_zz", 4, 2), ("R_x", 2),
                  ("ALLTOALL", "CZ")]
        specs += [("R_z", i) for i in range(N)] + [("CHAIN", "CPHASE"), ("RR_block", "R_yy")]
        return specs, N, None
    raise KeyError(name)


def spec_from_gate(g):
    """Reference-API gate object -> oracle spec, by public attributes only."""
    k = type(g).__name__
    if k in ("R_x", "R_y", "R_z", "negative_R_z", "I", "H", "X", "S", "T"):
        return (k, g.q_on)
    if k == "offset_R_z":
        return (k, g.q_on, g.offset)
    if k in ("fixed_R_y", "fixed_R_z"):
        return (k, g.q_on, g.theta)
    if k in ("CNOT", "CPHASE", "CZ", "sqrtiSWAP", "R_xx", "R_yy", "R_zz", "fSim", "fixed_fSim"):
        return (k, g.q1, g.q2)
    if k in ("CHAIN", "ALLTOALL"):
        return (k, g.entangler.__name__)
    if k == "RR_block":
        return (k, g.rotator.__name__)
    if k == "shared_parameter":
        return (k, [spec_from_gate(m) for m in g.layer], g.commute)
    raise KeyError(k)


def specs_from_circuit(qc):
    return [spec_from_gate(g) for g in qc.gates]
