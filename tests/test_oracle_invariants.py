"""The oracle has no reference numbers to be pinned against (the reference's tests
assert nothing and Julia is not installed), so it is validated by invariants of the scheme."""
import math

import numpy as np
import pytest

import problems as P
from oracle.oracle import Oracle, run_ssp33
from p2de_b200 import (EntropyStable, LowOrderPositivity, NoRHSLimiter, StandardDG, SubcellLimiter,
                       ZhangShuLimiter, primitive_to_conservative)


def make(problem, threads=2):
    param, rd, md, dd, bc, U0 = P.setup(problem)
    orc = Oracle(param, dd, bc, threads=threads)
    orc.set_state(U0)
    return param, md, dd, orc, U0


@pytest.mark.parametrize("limiter", [SubcellLimiter(), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
@pytest.mark.parametrize("N", [1, 3])
def test_free_stream_preserved(limiter, N):
    """Constant state: every RHS vanishes and no limiting happens."""
    param, ic, bc = P.vortex(N=N, K=(4, 3), limiter=limiter)
    const = lambda prm, x, y: primitive_to_conservative(prm.equation, (1.2 + 0 * x, 0.3 + 0 * x, -0.4 + 0 * x, 0.9 + 0 * x))
    param, md, dd, orc, U0 = make((param, const, bc))
    orc.rhs(0.0, 1e-2, 1)
    for f in ("rhsU", "rhsH", "rhsL"):
        assert np.abs(orc.field(f)).max() < 1e-12
    L = orc.field("L_local")[0] if limiter.code == 2 else orc.field("L")[0]
    assert (L == 1.0).all()


@pytest.mark.parametrize("rhs_type", [None, LowOrderPositivity(), EntropyStable(), StandardDG()],
                         ids=["limited", "low", "es", "stddg"])
def test_conservation_periodic(rhs_type):
    """sum_k sum_i wJ rhsU = 0 on a periodic mesh (check_conservation, dg/utils.jl:1-12)."""
    kw = {} if rhs_type is None else dict(rhs=rhs_type, limiter=NoRHSLimiter())
    param, md, dd, orc, U0 = make(P.kelvin_helmholtz(N=3, K=(6, 5), **kw))
    orc.rhs(0.0, 5e-4, 1)
    wJ = dd.ops.wq[None, :, None] * dd.geom.Jq[:, :, None]
    tot = (wJ * orc.field("rhsU")).sum((0, 1))
    scale = (wJ * np.abs(orc.field("rhsU"))).sum((0, 1)).max()
    assert np.abs(tot).max() < 1e-13 * max(scale, 1.0)


def test_conservation_over_steps_and_dt_rules():
    param, md, dd, orc, U0 = make(P.vortex(N=2, K=(6, 6), CFL=0.5, T=0.05, dt0=1e-2))
    c0 = orc.reduce(0)
    t, dth = run_ssp33(orc, 0.0, param.timestepping_param.T)
    assert abs(orc.reduce(0) - c0) < 1e-12 * abs(c0)
    assert abs(t - param.timestepping_param.T) < 1e-14          # last step is clipped to T - t (SSPRK33.jl:30)
    assert max(dth) <= 0.5 * 1e-2 * (1 + 1e-15)                 # dt <= CFL*dt0


def test_stage1_limiter_sees_the_cap_not_the_cfl_dt():
    """rhs!(::LimitedDG) computes the CFL dt but hands the ORIGINAL time_param to the limiter
    (rhs.jl:46,52): L_local depends on the dt passed in, dt_out does not."""
    param, md, dd, orc, U0 = make(P.sedov(N=2, K=(8, 8)))
    dta = orc.rhs(0.0, 5e-4, 1); La = orc.field("L_local")[0].copy()
    orc.set_state(U0)
    dtb = orc.rhs(0.0, 5e-2, 1); Lb = orc.field("L_local")[0].copy()
    assert dta == dtb
    assert (Lb <= La + 1e-15).all() and (Lb < La - 1e-3).any()


@pytest.mark.parametrize("problem", [P.sedov(N=3, K=(8, 8)), P.dmr(N=3, K=(24, 6)),
                                     P.sedov(N=2, K=(8, 8), limiter=ZhangShuLimiter())], ids=["sedov", "dmr", "sedov-zs"])
def test_positivity_preserved(problem):
    param, md, dd, orc, U0 = make(problem)
    t = 0.0
    for _ in range(15):
        t += orc.ssp33_step(t)
        assert orc.reduce(1) > 0 and orc.reduce(2) > 0
    L = orc.field("L_local") if param.rhs_limiter.code == 2 else orc.field("L")
    assert (L >= 0).all() and (L <= 1).all()


def test_symmetrised_coefficients_match_across_interfaces():
    """symmetrize_limiting_parameters! (subcell.jl:418-456): both sides of an interface agree."""
    param, md, dd, orc, U0 = make(P.sedov(N=2, K=(6, 6)))
    orc.rhs(0.0, 2e-2, 1)
    L = orc.field("L_local")[0]          # [K, d, idx]
    n = param.N + 1
    Kx = 6
    for k in range(dd.sizes.K):
        kr = (k // Kx) * Kx + (k % Kx + 1) % Kx
        for sj in range(n):
            assert L[k, 0, n + sj * (n + 1)] == L[kr, 0, 0 + sj * (n + 1)]
        kt = (k + Kx) % dd.sizes.K
        for si in range(n):
            assert L[k, 1, si + n * n] == L[kt, 1, si]


def test_vortex_convergence_order():
    """Isentropic vortex (test/test_smoke.jl:6-20), Zhang-Shu limiter inactive: L2 error drops at
    towards order N+1 under refinement (examples/convergence/isentropic-vortex-convergence.jl)."""
    errs = []
    for Kk in (12, 24):
        param, md, dd, orc, U0 = make(P.vortex(N=3, K=(Kk, Kk), limiter=ZhangShuLimiter(), CFL=0.5, T=0.2, dt0=1e-2), threads=4)
        t, _ = run_ssp33(orc, 0.0, param.timestepping_param.T)
        ex = np.stack(primitive_to_conservative(param.equation, P.vortex_exact(param.equation, md.xq, md.yq, t)), -1)
        wJ = dd.ops.wq[None, :] * dd.geom.Jq
        U = orc.get_state()
        errs.append(sum(math.sqrt((wJ * (U[..., c] - ex[..., c]) ** 2).sum()) / math.sqrt((wJ * ex[..., c] ** 2).sum()) for c in range(4)))
    assert math.log2(errs[0] / errs[1]) > 2.5
