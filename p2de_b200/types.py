"""Host-side mirror of the reference's option/parameter structs.

Same names, same fields, same defaults as `src/common/types/Solver.jl` so that a script
written against P2DE.jl reads the same here (the reference dispatches on these singleton
types at compile time; here they are mapped to the enums of `include/p2de_b200.h`).

    reference                                             here
    RHS / LowOrderPositivity / FluxDiffRHS / LimitedDG    Solver.jl:1-13
    ESLimitedLowOrderPos(...), EntropyStable(...), StandardDG     Solver.jl:24-36
    ZhangShuLimiter / SubcellLimiter(bound, shockcapture) Solver.jl:76-87
    Param (kwdef)                                         Solver.jl:151-170
    BCData / StateParam                                   StateParam.jl:1-11
    TimeParam                                             TimeParam.jl:1-6
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------- enums (ABI)
BASIS_LOBATTO, BASIS_GAUSS = 0, 1
RHS_LOW_ORDER_POSITIVITY, RHS_FLUX_DIFF, RHS_LIMITED_DG = 0, 1, 2
VOLFLUX_CHANDRASHEKAR, VOLFLUX_CENTRAL = 0, 1
SURFFLUX_CHANDRASHEKAR_PROJECTED, SURFFLUX_LF_NODAL, SURFFLUX_LF_PROJECTED = 0, 1, 2
PROJLIM_NONE, PROJLIM_NODEWISE = 0, 1
LIMITER_NONE, LIMITER_ZHANGSHU, LIMITER_SUBCELL = 0, 1, 2
(BOUND_POSITIVITY, BOUND_POS_MIN_ENTROPY, BOUND_POS_RELAXED_MIN_ENTROPY, BOUND_POS_CELL_ENTROPY,
 BOUND_POS_RELAXED_CELL_ENTROPY, BOUND_TVD, BOUND_TVD_MIN_ENTROPY, BOUND_TVD_RELAXED_MIN_ENTROPY,
 BOUND_TVD_CELL_ENTROPY, BOUND_TVD_RELAXED_CELL_ENTROPY) = range(10)
SHOCKCAPTURE_NONE, SHOCKCAPTURE_HENNEMANN = 0, 1

(FIELD_UQ, FIELD_RHSU, FIELD_RHSH, FIELD_RHSL, FIELD_L, FIELD_L_LOCAL, FIELD_THETA,
 FIELD_THETA_LOCAL, FIELD_RESW) = range(9)
REDUCE_CONSERVATION, REDUCE_MIN_RHO, REDUCE_MIN_RHOE = 0, 1, 2


# ----------------------------------------------------------------------------- flux types
@dataclass(frozen=True)
class ChandrashekarFlux:
    code: int = VOLFLUX_CHANDRASHEKAR


@dataclass(frozen=True)
class CentralFlux:
    code: int = VOLFLUX_CENTRAL


@dataclass(frozen=True)
class ChandrashekarOnProjectedVal:
    code: int = SURFFLUX_CHANDRASHEKAR_PROJECTED


@dataclass(frozen=True)
class LaxFriedrichsOnNodalVal:
    code: int = SURFFLUX_LF_NODAL


@dataclass(frozen=True)
class LaxFriedrichsOnProjectedVal:
    code: int = SURFFLUX_LF_PROJECTED


# ----------------------------------------------------------------------------- RHS types
@dataclass(frozen=True)
class LowOrderPositivity:
    surface_flux: Any = LaxFriedrichsOnNodalVal()
    code: int = RHS_LOW_ORDER_POSITIVITY


@dataclass(frozen=True)
class FluxDiffRHS:
    volume_flux: Any = ChandrashekarFlux()
    surface_flux: Any = LaxFriedrichsOnProjectedVal()
    code: int = RHS_FLUX_DIFF


@dataclass(frozen=True)
class LimitedDG:
    low_order_surface_flux: Any = LaxFriedrichsOnNodalVal()
    high_order_surface_flux: Any = LaxFriedrichsOnProjectedVal()
    high_order_volume_flux: Any = ChandrashekarFlux()
    code: int = RHS_LIMITED_DG


def EntropyStable(surface_flux=LaxFriedrichsOnProjectedVal()):
    return FluxDiffRHS(ChandrashekarFlux(), surface_flux)


def StandardDG():
    return FluxDiffRHS(CentralFlux(), LaxFriedrichsOnProjectedVal())


def ESLimitedLowOrderPos(low_order_surface_flux=LaxFriedrichsOnNodalVal(),
                         high_order_surface_flux=LaxFriedrichsOnProjectedVal()):
    return LimitedDG(low_order_surface_flux, high_order_surface_flux, ChandrashekarFlux())


def StdDGLimitedLowOrderPos(low_order_surface_flux=LaxFriedrichsOnNodalVal(),
                            high_order_surface_flux=LaxFriedrichsOnProjectedVal()):
    return LimitedDG(low_order_surface_flux, high_order_surface_flux, CentralFlux())


# ----------------------------------------------------------------------------- limiters
@dataclass(frozen=True)
class NoEntropyProjectionLimiter:
    code: int = PROJLIM_NONE


@dataclass(frozen=True)
class NodewiseScaledExtrapolation:
    code: int = PROJLIM_NODEWISE


@dataclass(frozen=True)
class PositivityBound:
    code: int = BOUND_POSITIVITY


@dataclass(frozen=True)
class PositivityAndMinEntropyBound:
    code: int = BOUND_POS_MIN_ENTROPY


@dataclass(frozen=True)
class PositivityAndRelaxedMinEntropyBound:
    code: int = BOUND_POS_RELAXED_MIN_ENTROPY


@dataclass(frozen=True)
class PositivityAndCellEntropyBound:
    code: int = BOUND_POS_CELL_ENTROPY


@dataclass(frozen=True)
class PositivityAndRelaxedCellEntropyBound:
    beta: float = 0.5
    code: int = BOUND_POS_RELAXED_CELL_ENTROPY


@dataclass(frozen=True)
class TVDBound:
    code: int = BOUND_TVD


@dataclass(frozen=True)
class TVDAndMinEntropyBound:
    code: int = BOUND_TVD_MIN_ENTROPY


@dataclass(frozen=True)
class TVDAndRelaxedMinEntropyBound:
    code: int = BOUND_TVD_RELAXED_MIN_ENTROPY


@dataclass(frozen=True)
class TVDAndCellEntropyBound:
    code: int = BOUND_TVD_CELL_ENTROPY


@dataclass(frozen=True)
class TVDAndRelaxedCellEntropyBound:
    beta: float = 0.5
    code: int = BOUND_TVD_RELAXED_CELL_ENTROPY


@dataclass(frozen=True)
class NoShockCapture:
    code: int = SHOCKCAPTURE_NONE


@dataclass(frozen=True)
class HennemannShockCapture:
    a: float = 0.5
    c: float = 1.8
    code: int = SHOCKCAPTURE_HENNEMANN


@dataclass(frozen=True)
class NoRHSLimiter:
    code: int = LIMITER_NONE


@dataclass(frozen=True)
class ZhangShuLimiter:
    shockcapture: Any = NoShockCapture()
    code: int = LIMITER_ZHANGSHU

    @property
    def bound(self):  # traits.jl: bound(::ZhangShuLimiter) = PositivityBound()
        return PositivityBound()


@dataclass(frozen=True)
class SubcellLimiter:
    bound: Any = PositivityBound()
    shockcapture: Any = NoShockCapture()
    code: int = LIMITER_SUBCELL


# ----------------------------------------------------------------------------- basis / equation
@dataclass(frozen=True)
class GaussCollocation:
    code: int = BASIS_GAUSS


@dataclass(frozen=True)
class LobattoCollocation:
    code: int = BASIS_LOBATTO


@dataclass(frozen=True)
class CompressibleEulerIdealGas:
    """CompressibleEulerIdealGas{Dim1|Dim2}(gamma)  (Solver.jl:110-119)."""
    dim: int
    gamma: float

    @property
    def Nc(self):
        return self.dim + 2


def get_gamma(equation):
    return equation.gamma


def num_components(equation):
    return equation.Nc


# ----------------------------------------------------------------------------- parameters
@dataclass(frozen=True)
class GlobalConstant:
    POSTOL: float
    ZEROTOL: float


@dataclass(frozen=True)
class TimesteppingParameter:
    T: float
    CFL: float
    dt0: float
    t0: float


@dataclass(frozen=True)
class PostprocessingParameter:
    output_interval: int


@dataclass(frozen=True)
class LimitingParameter:
    zeta: float
    eta: float


@dataclass(frozen=True)
class Param:
    N: int
    K: Any                     # Int in 1D, (Kx, Ky) in 2D
    xL: Any
    xR: Any
    global_constants: GlobalConstant
    timestepping_param: TimesteppingParameter
    limiting_param: LimitingParameter
    postprocessing_param: PostprocessingParameter
    equation: CompressibleEulerIdealGas
    approximation_basis: Any
    rhs: Any
    entropyproj_limiter: Any
    rhs_limiter: Any


def num_elements(param: Param) -> int:
    return int(np.prod(param.K))


@dataclass
class BCData:
    """BCData{Nc}(mapP, mapI, mapO, Ival): indices are 1-based linear indices into [Nfp,K]."""
    mapP: np.ndarray                 # int64 [K, Nfp] (C order == Julia [Nfp, K])
    mapI: np.ndarray
    mapO: np.ndarray
    Ival: np.ndarray                 # [nI, Nc]

    def __post_init__(self):
        self.mapP = np.ascontiguousarray(self.mapP, dtype=np.int64)
        self.mapI = np.ascontiguousarray(np.asarray(self.mapI, dtype=np.float64), dtype=np.int64).reshape(-1)
        self.mapO = np.ascontiguousarray(np.asarray(self.mapO, dtype=np.float64), dtype=np.int64).reshape(-1)
        Ival = np.asarray(self.Ival, dtype=np.float64)
        self.Ival = np.ascontiguousarray(Ival.reshape(len(self.mapI), -1) if Ival.size else Ival.reshape(0, 0))


@dataclass
class StateParam:
    bcdata: BCData


@dataclass
class TimeParam:
    t: float
    dt: float
    nstage: int
    timer: Any = None


@dataclass
class SizeData:
    K: int
    N1D: int
    Nd: int
    Nc: int
    Np: int
    Nq: int
    Nfp: int
    Nh: int
    Ns: int


@dataclass
class GeomData:
    J: np.ndarray      # [K, Nq]
    Jq: np.ndarray     # [K, Nq]
    GJh: Tuple[np.ndarray, ...]   # each [K, Nh]


@dataclass
class Operators:
    Srsh_db: Tuple[np.ndarray, ...]    # each (Nh, Nh) math-indexed [i, j]
    Srs0: Tuple[np.ndarray, ...]       # each (Nq, Nq) dense
    Srsh_nnz: Sequence[Tuple[int, int]]   # 1-based (i, j), i > j, reference order
    Srs0_nnz: Sequence[Tuple[int, int]]
    Brs: Tuple[np.ndarray, ...]        # each (Nfp,) diagonal
    Vh: np.ndarray
    MinvVhT: np.ndarray
    VDM_inv: np.ndarray
    Vq: np.ndarray
    Vf: np.ndarray
    Vf_low: np.ndarray
    Pq: np.ndarray
    MinvVfT: np.ndarray
    wq: np.ndarray
    q2fq: Sequence[Sequence[int]]      # 1-based
    fq2q: np.ndarray                   # 1-based int64


@dataclass
class Discretization:
    sizes: SizeData
    geom: GeomData
    ops: Operators


@dataclass
class MeshData:
    """The subset of StartUpDG.MeshData the reference's callbacks touch."""
    K: int
    xq: np.ndarray     # [K, Nq]
    yq: Any            # [K, Nq] | None
    xf: np.ndarray     # [K, Nfp]
    yf: Any
    mapM: np.ndarray   # [K, Nfp] 1-based
    mapP: np.ndarray   # [K, Nfp] 1-based (self on domain boundaries until make_periodic)
    mapB: np.ndarray   # 1-based linear indices of boundary face nodes
    J: np.ndarray
    rxJ: float
    sxJ: float
    ryJ: float
    syJ: float
    Kxy: Tuple[int, ...] = ()
    is_periodic: Tuple[bool, ...] = ()


@dataclass
class Solver:
    param: Param
    rd: Any
    md: MeshData
    discrete_data: Discretization


@dataclass
class DataHistory:
    Uhist: list = field(default_factory=list)
    Lhist: list = field(default_factory=list)
    thetahist: list = field(default_factory=list)
    thist: list = field(default_factory=list)
    dthist: list = field(default_factory=list)


@dataclass
class ErrorData:
    L1err: float
    L2err: float
    Linferr: float
