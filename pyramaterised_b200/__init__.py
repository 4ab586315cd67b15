"""pyramaterised_b200 -- a B200-native engine behind the pyramaterised API.

Same public names as the reference package (/root/reference/pyramaterised/__init__.py:1-4):
``PQC`` and every gate class at top level, plus the ``measure``, ``gates`` and ``templates``
sub-modules.  The numerical backend is libpqc_b200.so (hand-written sm_100a CUDA behind
the C ABI in include/pqc_b200.h); there is no QuTiP and no CPU fallback.
"""
from .circuit import *          # noqa: F401,F403
from . import measure           # noqa: F401
from . import gates             # noqa: F401
from . import templates         # noqa: F401
from . import engine            # noqa: F401
from . import qobj              # noqa: F401
from .qobj import State, PauliSum   # noqa: F401

__version__ = "0.1.0"
