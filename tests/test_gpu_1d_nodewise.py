"""GPU parity for NodewiseScaledExtrapolation on 1D Gauss nodes: the entropy-projection limiting
parameter theta per face node (src/dg/filter.jl:6-130: 21-step bisection on the blended extrapolation of the entropy
variables), the projection with theta (rhs.jl:84-94), the limited face matrix in assemble_rhs!
(flux_differencing.jl:288-319) and LaxFriedrichsOnProjectedVal's find_alpha on the limited face state
(low_order_graph_viscosity.jl:299-327), through the C ABI (csrc/kernels1d.cuh: theta_face1, project_face) against the oracle.

theta = 1 everywhere on the shock tubes' initial data (constant inside every element), so the comparisons start from a state
the oracle has advanced until the projection limiter bites.  Tolerances as in test_gpu_gauss.py: 1e-11 relative on one rhs!,
theta_local to 1e-9 with identical {theta == 1} sets."""
import numpy as np
import pytest

import problems as P
from p2de_b200 import (ESLimitedLowOrderPos, GaussCollocation, LaxFriedrichsOnProjectedVal, LobattoCollocation,
                       NodewiseScaledExtrapolation, SubcellLimiter, TimeParam, ZhangShuLimiter)
from p2de_b200 import types as T
from test_gpu_parity import make_pair, rel

pytestmark = pytest.mark.gpu

NW = dict(basis=GaussCollocation(), entropyproj_limiter=NodewiseScaledExtrapolation())
PROJ = ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), LaxFriedrichsOnProjectedVal())

CASES = {
    "sod-N3": lambda **kw: P.sod(N=3, K=40, **NW, **kw),
    "sod-N2": lambda **kw: P.sod(N=2, K=50, **NW, **kw),       # (the limiter bites intermittently here: steps 13-23, 37-48, ...)
    "shu-osher-N3": lambda **kw: P.shu_osher(N=3, K=64, **NW, **kw),
    "leblanc-N2": lambda **kw: P.leblanc(N=2, K=100, **NW, **kw),
}


DEVELOP = {"sod-N2": 15}     # oracle steps before the rhs comparison (default 30): the next stage evaluation has theta < 1


def developed_pair(problem, nsteps, keep_diagnostics=True):
    """(param, solver, st, orc, t): oracle advanced `nsteps`, both sides set to that state."""
    param, solver, st, orc, U0 = make_pair(problem, keep_diagnostics=keep_diagnostics)
    t = param.timestepping_param.t0
    for _ in range(nsteps):
        t += orc.ssp33_step(t)
    U = orc.get_state().copy()
    st.set_state(U)
    return param, solver, st, orc, t


def check_rhs(param, solver, st, orc, t, nstage, dt, rtol=1e-11):
    from p2de_b200.api import rhs
    dt_o = orc.rhs(t, dt, nstage)
    dt_g = rhs(st, solver, None, TimeParam(t=t, dt=dt, nstage=nstage))
    assert abs(dt_g - dt_o) <= 1e-12 * abs(dt_o), (dt_g, dt_o)
    pre = st.preallocation
    tg = pre.theta_local.reshape(3, -1)[nstage - 1]
    to = orc.field("theta_local").reshape(3, -1)[nstage - 1]
    assert np.abs(tg - to).max() < 1e-9
    assert np.array_equal(tg == 1.0, to == 1.0)
    assert np.abs(pre.theta.reshape(3, -1)[nstage - 1] - orc.field("theta").reshape(3, -1)[nstage - 1]).max() < 1e-9
    for f in ("rhsL", "rhsH", "rhsU"):
        assert rel(getattr(pre, f), orc.field(f)) < rtol, f
    if param.rhs_limiter.code == T.LIMITER_SUBCELL:
        Lg, Lo = pre.L_local[nstage - 1], orc.field("L_local")[nstage - 1]
        assert np.abs(Lg - Lo).max() < 1e-10
    else:
        assert np.abs(pre.L[nstage - 1] - orc.field("L")[nstage - 1]).max() < 1e-10
    return to


@pytest.mark.parametrize("limiter", [SubcellLimiter(), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_1d_nodewise_rhs_per_stage(name, limiter):
    """One rhs! per stage index on a developed shock-tube state; the projection limiter bites (theta < 1 somewhere)."""
    bites = False
    for nstage in (1, 2, 3):
        param, solver, st, orc, t = developed_pair(CASES[name](limiter=limiter), DEVELOP.get(name, 30))
        tp = param.timestepping_param
        th = check_rhs(param, solver, st, orc, t, nstage, tp.CFL * tp.dt0)
        bites = bites or (th < 1.0).any()
        st.close()
    assert bites


@pytest.mark.parametrize("name", ["sod-N3", "leblanc-N2"])
def test_1d_nodewise_projected_low_order_flux(name):
    """LaxFriedrichsOnProjectedVal for both surface fluxes: the low-order flux and find_alpha see the limited face state."""
    for nstage in (1, 2):
        param, solver, st, orc, t = developed_pair(CASES[name](rhs=PROJ), 30)
        tp = param.timestepping_param
        check_rhs(param, solver, st, orc, t, nstage, tp.CFL * tp.dt0)
        st.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_1d_nodewise_ssp33_steps(name):
    """40 SSP33! steps from the initial data: states within 1e-8 relative, positive, theta of the last step agrees."""
    param, solver, st, orc, U0 = make_pair(CASES[name](), keep_diagnostics=False)
    t_o = t_g = param.timestepping_param.t0
    seen = False
    for _ in range(40):
        dto = orc.ssp33_step(t_o); t_o += dto
        dtg = st.ssp33_step(t_g); t_g += dtg
        assert abs(dtg - dto) <= 1e-10 * dto
        seen = seen or (orc.field("theta_local") < 1.0).any()
    Ug, Uo = st.preallocation.Uq, orc.get_state()
    assert seen
    assert rel(Ug, Uo) < 1e-8
    assert (Ug[..., 0] > 0).all() and st.reduce(T.REDUCE_MIN_RHOE) > 0
    assert abs(st.reduce(T.REDUCE_CONSERVATION) - orc.reduce(0)) < 1e-10 * abs(orc.reduce(0))
    tg = st.preallocation.theta_local.reshape(-1)
    assert np.abs(tg - orc.field("theta_local").reshape(-1)).max() < 1e-6
    st.close()


def test_1d_nodewise_on_lobatto_nodes_is_the_identity():
    """filter.jl:18-20: theta_local = 1 on Lobatto nodes, theta is never written; results equal NoEntropyProjectionLimiter."""
    from p2de_b200.api import rhs
    out = []
    for kw in (dict(entropyproj_limiter=NodewiseScaledExtrapolation()), dict()):
        param, solver, st, orc, t = developed_pair(P.sod(N=3, K=40, basis=LobattoCollocation(), **kw), 20)
        tp = param.timestepping_param
        rhs(st, solver, None, TimeParam(t=t, dt=tp.CFL * tp.dt0, nstage=1))
        orc.rhs(t, tp.CFL * tp.dt0, 1)
        pre = st.preallocation
        if kw:
            assert (pre.theta_local.reshape(3, -1)[0] == 1.0).all() and (pre.theta == 0.0).all()
            assert np.array_equal(pre.theta_local.reshape(3, -1)[0], orc.field("theta_local").reshape(3, -1)[0])
        assert rel(pre.rhsU, orc.field("rhsU")) < 1e-12
        out.append(pre.rhsU.copy())
        st.close()
    assert np.array_equal(out[0], out[1])
