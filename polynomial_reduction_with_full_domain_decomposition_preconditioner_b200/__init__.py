"""
polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 (prfdd_b200)

B200-native rebuild of the PR-FDD preconditioned Krylov hot path.  The product is the C-ABI shared
library libprfdd_b200.so (hand-written sm_100a CUDA + C++ host classes mirroring the reference's
Domain / Subdomain / CSR_Matrix / Math / Element / Timer); this Python package is only the ctypes
binding used by tests and bench.py.  There is no CPU fallback: importing `capi` raises if the library
has not been built.
"""
from . import capi  # noqa: F401
from .capi import Solver, Options, lib, mesh_generate_box, ladder  # noqa: F401
