"""The reference-side binding of INTEGRATION.md (integration/pqc_b200_binding.py) against the
UNMODIFIED reference: `lower` applied to circuits built from /root/reference's own gate objects
must give exactly the op list pyramaterised_b200's classes give for the same builder.  The
reference only exists in the authoring container (its QuTiP calls run on oracle/qutip_lite, as
for the golden fixtures); elsewhere these tests skip."""
import importlib
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pyramaterised")),
                                reason="the reference tree is only present in the authoring container")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    from oracle import qutip_lite
    saved = {k: sys.modules.get(k) for k in ("qutip", "qutip.qip", "qutip.qip.operations",
                                             "qutip.states", "qutip.random_objects")}
    qutip_lite.install_as_qutip()
    sys.path.insert(0, REF)
    try:
        yield importlib.import_module("pyramaterised")
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _binding():
    sys.path.insert(0, os.path.join(ROOT, "integration"))
    try:
        return importlib.import_module("pqc_b200_binding")
    finally:
        sys.path.pop(0)


def test_lower_of_reference_circuits_equals_package_lowering(ref):
    import cases
    import pyramaterised_b200 as pyqc
    b = _binding()
    for name, (builder, *_rest) in sorted(cases.CASES.items()):
        theirs = builder(ref)
        ours = builder(pyqc)
        ang = list(cases.case_angles(name, ours, 1)[0])
        theirs.set_params(ang)            # fixed gates and frozen angles are read from the objects
        ours.set_params(ang)
        lo_ref, lo_own = b.lower(theirs), ours.lower()
        assert len(lo_ref) == len(lo_own), name
        for x, y in zip(lo_ref, lo_own):
            assert tuple(x[:6]) == tuple(y[:6]), (name, x, y)
            assert np.allclose(x[6:], y[6:], rtol=0, atol=0), (name, x, y)


def test_binding_struct_matches_header():
    b = _binding()
    hdr = open(os.path.join(ROOT, "include", "pqc_b200.h")).read()
    body = hdr[hdr.index("enum pqc_opcode {"):hdr.index("PQC_OP__COUNT")]
    names = [ln.split("=")[0].strip() for ln in body.splitlines() if ln.strip().startswith("PQC_OP_")]
    want = ["RX", "RY", "RZ", "H", "X", "S", "T", "CNOT", "CZ", "SQRTISWAP", "RXX", "RYY", "RZZ", "FSIM",
            "FIXED_FSIM", "IDENT"]
    assert names == ["PQC_OP_" + w for w in want]
    assert [getattr(b, w) for w in want] == list(range(16))
    assert [f[0] for f in b.PqcOp._fields_] == ["kind", "q0", "q1", "param", "param2", "group", "scale",
                                                "offset"]
