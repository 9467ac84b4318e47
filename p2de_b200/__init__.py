"""p2de_b200 — B200-native replacement of the per-stage DG RHS + limiter hot path of P2DE.jl.

Host-side mirror of the reference's operator surface (`initialize_DG`, `rhs!`,
`apply_rhs_limiter!`, `SSP33!`, the `Param`/`Solver`/`State`/`BCData` structs) on top of
the C ABI of `libp2de_b200.so` (include/p2de_b200.h).  There is no CPU fallback.
"""
from .types import *  # noqa: F401,F403
from .init import (RefElemData, build_operators, element_nodes, gauss_quad, gauss_lobatto_quad,  # noqa: F401
                   initialize_data, light_mesh, make_periodic,
                   primitive_to_conservative, sample_initial_condition, structured_mapP)
from .api import (SSP33, State, apply_rhs_limiter, calculate_error, check_conservation,  # noqa: F401,E402
                  initialize_DG, rhs)
