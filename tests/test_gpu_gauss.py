"""GPU parity for row 8f-1: 2D Gauss collocation, with and without NodewiseScaledExtrapolation
(src/dg/filter.jl, rhs.jl:97-133, flux_differencing.jl:164-361 with the hybridized face-volume pairs and the
limited Vf in assemble_rhs!, low_order_graph_viscosity.jl:249-327), through the C ABI against the oracle.
Kernels: gauss_project_kernel (csrc/gauss.cuh) + the generic stage kernel (csrc/kernels2d.cuh, A.gauss).

Tolerance: 1e-11 relative on one rhs! (pow / exp / log of the entropy-variable round trip differ between
libm and CUDA by an ulp or two and are amplified by 1/(gamma-1) exponents); theta_local to 1e-9 absolute
with identical {theta == 1} sets."""
import numpy as np
import pytest

import problems as P
from p2de_b200 import (ESLimitedLowOrderPos, EntropyStable, GaussCollocation, LaxFriedrichsOnNodalVal,
                       LaxFriedrichsOnProjectedVal, ChandrashekarOnProjectedVal, LowOrderPositivity,
                       NodewiseScaledExtrapolation, NoRHSLimiter, SubcellLimiter, TimeParam, ZhangShuLimiter)
from p2de_b200 import types as T
from test_gpu_parity import make_pair, rel

pytestmark = pytest.mark.gpu

RHS = ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), LaxFriedrichsOnProjectedVal())
GAUSS = dict(basis=GaussCollocation(), rhs=RHS)
NODEWISE = dict(basis=GaussCollocation(), entropyproj_limiter=NodewiseScaledExtrapolation(), rhs=RHS)
RTOL = 1e-11


def check(problem, nstage=1, dt=None, rtol=RTOL):
    from p2de_b200.api import rhs
    param, solver, st, orc, U0 = make_pair(problem)
    tp = param.timestepping_param
    dt = tp.CFL * tp.dt0 if dt is None else dt
    dt_o = orc.rhs(tp.t0, dt, nstage)
    dt_g = rhs(st, solver, None, TimeParam(t=tp.t0, dt=dt, nstage=nstage))
    assert abs(dt_g - dt_o) <= 1e-12 * abs(dt_o), (dt_g, dt_o)
    pre = st.preallocation
    code = param.rhs.code
    if code != T.RHS_FLUX_DIFF:
        assert rel(pre.rhsL, orc.field("rhsL")) < rtol
    if code != T.RHS_LOW_ORDER_POSITIVITY:
        assert rel(pre.rhsH, orc.field("rhsH")) < rtol
    assert rel(pre.rhsU, orc.field("rhsU")) < rtol
    out = {}
    if param.entropyproj_limiter.code == T.PROJLIM_NODEWISE:
        sz = solver.discrete_data.sizes
        tg = pre.theta_local.reshape(3, -1)[nstage - 1]
        to = orc.field("theta_local").reshape(3, -1)[nstage - 1]
        assert np.abs(tg - to).max() < 1e-9
        assert np.array_equal(tg == 1.0, to == 1.0)
        assert np.abs(pre.theta.reshape(3, -1)[nstage - 1] - orc.field("theta").reshape(3, -1)[nstage - 1]).max() < 1e-9
        out["theta"] = to
    if code == T.RHS_LIMITED_DG and param.rhs_limiter.code == T.LIMITER_SUBCELL:
        Lg, Lo = pre.L_local[nstage - 1], orc.field("L_local")[nstage - 1]
        assert np.abs(Lg - Lo).max() < 1e-10
        out["L"] = Lo
    return out


@pytest.mark.parametrize("N", [1, 2, 3, 4])
@pytest.mark.parametrize("limiter", [SubcellLimiter(), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
def test_gauss_rhs_without_projection_limiter(N, limiter):
    """NoEntropyProjectionLimiter on data whose unlimited projection stays positive (on the strong test vortex
    it does not: that is what NodewiseScaledExtrapolation is for)."""
    # dt = 1e-4: at the default cap the N=1 low-order update turns one density negative, and the coefficient of
    # a subcell face with f_H - f_L = rounding noise is then decided by the sign of that noise (0 or 1) in any build
    for nstage in (1, 2, 3):
        check(P.kelvin_helmholtz(N=N, K=(5, 6), limiter=limiter, **GAUSS), nstage=nstage, dt=1e-4)


@pytest.mark.parametrize("N", [1, 2, 3, 4])
@pytest.mark.parametrize("limiter", [SubcellLimiter(), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
def test_gauss_nodewise_smoke_vortex_rhs(N, limiter):
    """test/test_smoke.jl:44-67 scenario on Gauss nodes with NodewiseScaledExtrapolation."""
    for nstage in (1, 2, 3):
        check(P.vortex(N=N, K=(5, 5), limiter=limiter, **NODEWISE), nstage=nstage)


@pytest.mark.parametrize("N", [2, 3])
def test_gauss_nodewise_rhs_with_active_theta(N):
    """The strong test vortex on a coarse mesh trips the bounds of filter.jl:84-98 at some face nodes."""
    out = check(P.vortex(N=N, K=(8, 8), **NODEWISE))
    assert (out["theta"] < 1.0).any()
    check(P.kelvin_helmholtz(N=N, K=(6, 6), **NODEWISE), nstage=2)


@pytest.mark.parametrize("rhs_type", [
    LowOrderPositivity(LaxFriedrichsOnProjectedVal()),
    LowOrderPositivity(LaxFriedrichsOnNodalVal()),
    EntropyStable(LaxFriedrichsOnProjectedVal()),
    EntropyStable(ChandrashekarOnProjectedVal()),
    ESLimitedLowOrderPos(LaxFriedrichsOnNodalVal(), LaxFriedrichsOnProjectedVal()),
    ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), ChandrashekarOnProjectedVal()),
], ids=["low-proj", "low-nodal", "es-lf", "es-chand", "lim-lownodal", "lim-chand"])
def test_gauss_rhs_types(rhs_type):
    lim = NoRHSLimiter() if rhs_type.code != T.RHS_LIMITED_DG else SubcellLimiter()
    check(P.kelvin_helmholtz(N=3, K=(6, 6), basis=GaussCollocation(), rhs=rhs_type, limiter=lim))


def test_gauss_dmr_boundary_conditions_and_limiter():
    out = check(P.dmr(N=3, K=(16, 4), **NODEWISE), dt=5e-4)
    out = check(P.sedov(N=3, K=(8, 8), **NODEWISE), dt=2e-2)
    assert (out["L"] < 1.0).any()


def test_gauss_kelvin_helmholtz_steps_match_oracle():
    """examples/2D/kelvin-helmholtz.jl:44-55 configuration, a few SSP-RK3 steps."""
    param, solver, st, orc, U0 = make_pair(P.kelvin_helmholtz(N=3, K=(8, 8), **NODEWISE), keep_diagnostics=False)
    t = 0.0
    for _ in range(5):
        dto = orc.ssp33_step(t)
        dtg = st.ssp33_step(t)
        assert abs(dtg - dto) <= 1e-11 * dto
        t += dto
    assert rel(st.preallocation.Uq, orc.get_state()) < 1e-9
