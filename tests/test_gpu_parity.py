"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances: one rhs! evaluation <= 1e-12 relative (max-norm scaled by the field
max); limiting coefficients <= 1e-12 absolute and identical {l == 1} sets; sign(rho), sign(rho e)
identical; after T steps a stated, problem-dependent tolerance.
"""
import numpy as np
import pytest

import problems as P
from p2de_b200 import (EntropyStable, ESLimitedLowOrderPos, LaxFriedrichsOnNodalVal, LaxFriedrichsOnProjectedVal,
                       ChandrashekarOnProjectedVal, LowOrderPositivity, PositivityBound, StandardDG,
                       StdDGLimitedLowOrderPos, SubcellLimiter, TimeParam, ZhangShuLimiter, NoRHSLimiter)
from p2de_b200 import types as T

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


ROUNDTRIP = False   # default build option: LGL face states are the nodal states


def make_pair(problem, keep_diagnostics=True, roundtrip=None):
    from oracle.oracle import Oracle
    from p2de_b200.api import State
    from p2de_b200.types import Solver
    param, rd, md, dd, bc, U0 = P.setup(problem)
    solver = Solver(param=param, rd=rd, md=md, discrete_data=dd)
    st = State(solver, bc, keep_diagnostics=keep_diagnostics,
               lgl_projection_roundtrip=ROUNDTRIP if roundtrip is None else roundtrip)
    st.set_state(U0)
    orc = Oracle(param, dd, bc)
    orc.set_state(U0)
    return param, solver, st, orc, U0


def check_rhs(problem, nstage=1, dt=None, roundtrip=None, RTOL=RTOL):
    from p2de_b200.api import rhs
    param, solver, st, orc, U0 = make_pair(problem, roundtrip=roundtrip)
    tp = param.timestepping_param
    dt = tp.CFL * tp.dt0 if dt is None else dt
    dt_o = orc.rhs(tp.t0, dt, nstage)
    dt_g = rhs(st, solver, None, TimeParam(t=tp.t0, dt=dt, nstage=nstage))
    assert abs(dt_g - dt_o) <= 1e-13 * abs(dt_o), (dt_g, dt_o)
    pre = st.preallocation
    assert rel(pre.rhsU, orc.field("rhsU")) < RTOL
    code = param.rhs.code
    if code != T.RHS_FLUX_DIFF:
        assert rel(pre.rhsL, orc.field("rhsL")) < RTOL
    if code != T.RHS_LOW_ORDER_POSITIVITY:
        assert rel(pre.rhsH, orc.field("rhsH")) < RTOL
    if code == T.RHS_LIMITED_DG:
        if param.rhs_limiter.code == T.LIMITER_SUBCELL:
            Lg, Lo = pre.L_local[nstage - 1], orc.field("L_local")[nstage - 1]
            assert np.abs(Lg - Lo).max() < 1e-12
            assert np.array_equal(Lg == 1.0, Lo == 1.0)
            return Lo
        Lg, Lo = pre.L[nstage - 1], orc.field("L")[nstage - 1]
        assert np.abs(Lg - Lo).max() < 1e-12
        assert np.array_equal(Lg == 1.0, Lo == 1.0)
        return Lo


@pytest.mark.parametrize("N", [1, 2, 3, 4])
@pytest.mark.parametrize("limiter", [SubcellLimiter(bound=PositivityBound()), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
def test_smoke_vortex_rhs(N, limiter):
    """test/test_smoke.jl:44-67 scenario, one rhs! evaluation per stage index."""
    for nstage in (1, 2, 3):
        check_rhs(P.vortex(N=N, K=(5, 5), limiter=limiter), nstage=nstage)


@pytest.mark.parametrize("N", [2, 3])
def test_limiter_active_rhs(N):
    """Large dt on a blast wave: the limiter must actually bite and still match."""
    L = check_rhs(P.sedov(N=N, K=(8, 8)), dt=2e-2)
    assert (L < 1.0).any() and (L < 0.5).any()
    L = check_rhs(P.sedov(N=N, K=(8, 8), limiter=ZhangShuLimiter()), dt=2e-2)
    assert (L < 1.0).any()


@pytest.mark.parametrize("rhs_type", [
    LowOrderPositivity(LaxFriedrichsOnNodalVal()),
    LowOrderPositivity(LaxFriedrichsOnProjectedVal()),
    EntropyStable(LaxFriedrichsOnProjectedVal()),
    EntropyStable(ChandrashekarOnProjectedVal()),
    StandardDG(),
    ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), LaxFriedrichsOnProjectedVal()),
    ESLimitedLowOrderPos(LaxFriedrichsOnNodalVal(), ChandrashekarOnProjectedVal()),
    StdDGLimitedLowOrderPos(),
], ids=["low-nodal", "low-proj", "es-lf", "es-chand", "stddg", "lim-lowproj", "lim-chand", "lim-central"])
def test_rhs_types(rhs_type):
    lim = NoRHSLimiter() if rhs_type.code != T.RHS_LIMITED_DG else SubcellLimiter()
    check_rhs(P.kelvin_helmholtz(N=3, K=(6, 6), rhs=rhs_type, limiter=lim))


@pytest.mark.parametrize("N", [1, 3])
def test_faithful_projection_roundtrip_option(N):
    """cfg.lgl_projection_roundtrip=1 evaluates u(v(U)) at the face nodes exactly like rhs.jl:84-94."""
    check_rhs(P.vortex(N=N, K=(5, 5)), roundtrip=True)
    check_rhs(P.sedov(N=N, K=(6, 6), limiter=ZhangShuLimiter()), dt=2e-2, roundtrip=True)
    check_rhs(P.dmr(N=N, K=(8, 4)), dt=5e-4, roundtrip=True)


def test_dmr_boundary_conditions_rhs():
    """Inflow (mapI/Ival) + outflow (mapO) faces through the reference's BC surface."""
    L = check_rhs(P.dmr(N=3, K=(16, 4)), dt=5e-4)
    check_rhs(P.dmr(N=2, K=(12, 4), limiter=ZhangShuLimiter()), dt=5e-4)
    assert (L < 1.0).any()


def run_both(problem, nsteps):
    param, solver, st, orc, U0 = make_pair(problem, keep_diagnostics=False)
    t_o = t_g = param.timestepping_param.t0
    for _ in range(nsteps):
        dto = orc.ssp33_step(t_o); t_o += dto
        dtg = st.ssp33_step(t_g); t_g += dtg
        assert abs(dtg - dto) <= 1e-10 * dto
    return param, st.preallocation.Uq, orc.get_state(), st, orc


@pytest.mark.parametrize("limiter", [SubcellLimiter(), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
def test_ssp33_vortex_steps(limiter):
    """10 SSP33! steps on the smooth vortex: <= 1e-9 relative."""
    param, Ug, Uo, st, orc = run_both(P.vortex(N=3, K=(8, 8), limiter=limiter, CFL=0.5, T=10.0), 10)
    assert rel(Ug, Uo) < 1e-9
    assert abs(st.reduce(T.REDUCE_CONSERVATION) - orc.reduce(0)) < 1e-10 * abs(orc.reduce(0))


def test_ssp33_dmr_steps_positivity():
    """20 steps of the DMR-type data: positivity signs identical, states within 1e-7 relative
    (the limiter's thresholds amplify last-bit differences across a Mach-10 shock)."""
    param, Ug, Uo, st, orc = run_both(P.dmr(N=3, K=(32, 8)), 20)
    assert rel(Ug, Uo) < 1e-7
    rhoe = lambda U: U[..., 3] - 0.5 * (U[..., 1] ** 2 + U[..., 2] ** 2) / U[..., 0]
    assert np.array_equal(np.sign(Ug[..., 0]), np.sign(Uo[..., 0])) and (Ug[..., 0] > 0).all()
    assert np.array_equal(np.sign(rhoe(Ug)), np.sign(rhoe(Uo))) and (rhoe(Ug) > 0).all()
    assert st.reduce(T.REDUCE_MIN_RHO) > 0 and st.reduce(T.REDUCE_MIN_RHOE) > 0


def test_generic_mapP_path_matches_structured():
    """A mapP that is not the structured pattern (elements renumbered) goes through the generic
    gather tables and must give the same answer as the oracle."""
    from oracle.oracle import Oracle
    from p2de_b200.api import State, rhs
    from p2de_b200.types import Solver, BCData
    param, rd, md, dd, bc, U0 = P.setup(P.vortex(N=2, K=(4, 4)))
    K, Nfp = bc.mapP.shape
    rng = np.random.default_rng(0)
    perm = rng.permutation(K)                 # new element index -> old element index
    inv = np.argsort(perm)
    old = bc.mapP[perm] - 1                   # [new k, f] -> old linear index
    mapP_new = inv[old // Nfp] * Nfp + old % Nfp + 1
    bc2 = BCData(mapP_new, [], [], [])
    solver = Solver(param=param, rd=rd, md=md, discrete_data=dd)
    st = State(solver, bc2, keep_diagnostics=True)
    st.set_state(U0[perm])
    orc = Oracle(param, dd, bc2); orc.set_state(U0[perm])
    dt = 5e-3
    dto = orc.rhs(0.0, dt, 1)
    dtg = rhs(st, solver, None, TimeParam(t=0.0, dt=dt, nstage=1))
    assert abs(dtg - dto) <= 1e-13 * dto
    assert rel(st.preallocation.rhsU, orc.field("rhsU")) < RTOL


# ---- 1D path (BASELINE.json configs 1-2) -----------------------------------
from p2de_b200 import GaussCollocation, LobattoCollocation  # noqa: E402


@pytest.mark.parametrize("basis", [LobattoCollocation(), GaussCollocation()], ids=["lgl", "gauss"])
@pytest.mark.parametrize("limiter", [SubcellLimiter(), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
@pytest.mark.parametrize("N", [1, 3])
def test_1d_sod_rhs(N, limiter, basis):
    """1D Sod tube, inflow/outflow BCs: one rhs! per stage index (large dt so the limiter bites)."""
    for nstage in (1, 2):
        check_rhs(P.sod(N=N, K=40, limiter=limiter, basis=basis), nstage=nstage, dt=5e-3, roundtrip=True)
    check_rhs(P.sod(N=N, K=40, limiter=limiter, basis=basis), dt=5e-3)


@pytest.mark.parametrize("rhs_type", [LowOrderPositivity(LaxFriedrichsOnProjectedVal()), EntropyStable(ChandrashekarOnProjectedVal()),
                                      StandardDG(), ESLimitedLowOrderPos(LaxFriedrichsOnProjectedVal(), ChandrashekarOnProjectedVal())],
                         ids=["low-proj", "es-chand", "stddg", "lim-proj-chand"])
def test_1d_rhs_types_periodic(rhs_type):
    lim = NoRHSLimiter() if rhs_type.code != T.RHS_LIMITED_DG else SubcellLimiter()
    # Gauss nodes next to a face node give |f| ~ 1e-3 in logmean's -da/(logL - logR): its relative
    # rounding error ~1e-16/|f| is amplified by S/(wJ); the oracle itself moves by 1.5e-11 on this
    # case when compiled with FMA contraction, so 1e-12 is not meaningful here.
    check_rhs(P.density_wave_1d(N=3, K=12, rhs=rhs_type, limiter=lim, basis=GaussCollocation()), roundtrip=True, RTOL=2e-10)
    check_rhs(P.density_wave_1d(N=2, K=12, rhs=rhs_type, limiter=lim), roundtrip=True)


@pytest.mark.parametrize("problem", [P.sod(N=3, K=50), P.sod(N=3, K=50, basis=GaussCollocation(), limiter=ZhangShuLimiter()),
                                     P.shu_osher(N=3, K=64)], ids=["sod-lgl-subcell", "sod-gauss-zs", "shu-osher"])
def test_1d_ssp33_steps(problem):
    """30 SSP33! steps of the 1D shock tubes: positivity and <= 1e-8 relative agreement."""
    param, Ug, Uo, st, orc = run_both(problem, 30)
    assert rel(Ug, Uo) < 1e-8
    assert (Ug[..., 0] > 0).all() and st.reduce(T.REDUCE_MIN_RHOE) > 0
    assert abs(st.reduce(T.REDUCE_CONSERVATION) - orc.reduce(0)) < 1e-10 * abs(orc.reduce(0))


# ---- the other limiter variants of the reference's smoke test (test/test_smoke.jl:44-52)
from p2de_b200 import (HennemannShockCapture, PositivityAndMinEntropyBound,  # noqa: E402
                       PositivityAndRelaxedMinEntropyBound)

VARIANTS = {
    "subcell-hennemann": SubcellLimiter(bound=PositivityBound(), shockcapture=HennemannShockCapture()),
    "subcell-minentropy": SubcellLimiter(bound=PositivityAndMinEntropyBound()),
    "subcell-relaxed-minentropy": SubcellLimiter(bound=PositivityAndRelaxedMinEntropyBound()),
    "zhangshu-hennemann": ZhangShuLimiter(shockcapture=HennemannShockCapture()),
}


@pytest.mark.parametrize("N", [1, 2, 3, 4])
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_smoke_limiter_variants_rhs(N, variant):
    if "minentropy" in variant:
        # The smoke scenario is ISENTROPIC: s_modified = 2.5 +- a few ulp everywhere, so the test
        # s(U + l P) >= min_stencil(s) - POSTOL inside the 21-step bisection (limiter_utils.jl:42-50)
        # is decided by rounding noise and no two builds (nor two thread counts of the reference)
        # agree bit for bit.  Checked loosely here, strictly on non-isentropic data below.
        from p2de_b200.api import rhs
        param, solver, st, orc, U0 = make_pair(P.vortex(N=N, K=(5, 5), limiter=VARIANTS[variant]))
        tp = param.timestepping_param
        orc.rhs(tp.t0, tp.CFL * tp.dt0, 1)
        rhs(st, solver, None, TimeParam(t=tp.t0, dt=tp.CFL * tp.dt0, nstage=1))
        Lg, Lo = st.preallocation.L_local[0], orc.field("L_local")[0]
        assert (np.abs(Lg - Lo) > 1e-9).mean() < 0.25 and (Lg >= 0).all() and (Lg <= 1).all()
        assert rel(st.preallocation.rhsU, orc.field("rhsU")) < 1e-2
        assert rel(st.preallocation.rhsL, orc.field("rhsL")) < RTOL
        assert rel(st.preallocation.rhsH, orc.field("rhsH")) < RTOL
        return
    for nstage in (1, 2):
        check_rhs(P.vortex(N=N, K=(5, 5), limiter=VARIANTS[variant]), nstage=nstage, RTOL=1e-11)


@pytest.mark.parametrize("N", [2, 3])
@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_limiter_variants_nonisentropic_rhs(N, variant):
    check_rhs(P.kelvin_helmholtz(N=N, K=(6, 6), limiter=VARIANTS[variant]), dt=5e-3, RTOL=1e-11)


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_limiter_variants_on_shocks(variant):
    """Blast wave with a large dt: shock capturing / entropy bounds actually bite."""
    L = check_rhs(P.sedov(N=3, K=(8, 8), limiter=VARIANTS[variant]), dt=2e-2, RTOL=1e-11)
    assert (L < 1.0).any()
    param, Ug, Uo, st, orc = run_both(P.sedov(N=2, K=(8, 8), limiter=VARIANTS[variant]), 8)
    assert rel(Ug, Uo) < 1e-7


def test_direct_schedule_matches_two_kernel_schedule(monkeypatch):
    """The default schedule of the subcell family (stage 1 writes W = U + cap rhsU, stages 2/3 exchange U/2 + dt rhsL and
    write the new state from the stage kernel; stage_subcell.cuh) against P2DE_NO_DIRECT=1, which runs the run-time kernel
    (plain shares, rhsU) followed by the dense update kernel after every stage, and P2DE_NO_DEFER=1, which materialises U1
    with the axpy kernel.  Same scheme, different association of the same sums: agreement to rounding."""
    from p2de_b200.api import State
    from p2de_b200.types import Solver
    for prob in (P.dmr(N=3, K=(24, 16)), P.sedov(N=3, K=(12, 12)), P.dmr(N=3, K=(64, 8))):
        param, rd, md, dd, bc, U0 = P.setup(prob)
        out = []
        for env in ({}, {"P2DE_NO_DIRECT": "1"}, {"P2DE_NO_DEFER": "1"}):
            for k in ("P2DE_NO_DIRECT", "P2DE_NO_DEFER"):
                monkeypatch.setenv(k, env.get(k, "0"))
            st = State(Solver(param=param, rd=rd, md=md, discrete_data=dd), bc)
            st.set_state(U0)
            t, dts = 0.0, []
            for _ in range(6):
                dt = st.ssp33_step(t); t += dt; dts.append(dt)
            out.append((st.preallocation.Uq, dts))
            st.close()
        for U, dts in out[1:]:
            assert np.allclose(dts, out[0][1], rtol=1e-12, atol=0)
            assert rel(U, out[0][0]) < 1e-11


# ---- BASELINE.json config 2 on the GPU: Shu-Osher to T = 1.8 against the reference's own WENO5 data, and the Leblanc tube
#      (examples/convergence/leblanc-convergence.jl) against its exact solution through calculate_error
def test_1d_shu_osher_gpu_matches_reference_weno5_data():
    """examples/1D/shu-osher.jl:70-74 overlays data/weno5_shuosher.mat on its plot; here the GPU run to T = 1.8 (N=3, K=128) is
    held to the same number as the oracle (tests/test_oracle_1d.py): relative L1 density difference below 1.2 %."""
    import os
    from p2de_b200.api import State, SSP33
    from p2de_b200.types import Solver
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "weno5_shuosher_sub.npz"))
    param, rd, md, dd, bc, U0 = P.setup(P.shu_osher(N=3, K=128))
    solver = Solver(param=param, rd=rd, md=md, discrete_data=dd)
    st = State(solver, bc)
    st.set_state(U0)
    hist = SSP33(st, solver, None)
    assert abs(hist.thist[-1] - 1.8) < 1e-10
    U = st.preallocation.Uq
    x, rho = md.xq.reshape(-1), U[..., 0].reshape(-1)
    ref = np.interp(x, g["x"], g["rho"])
    w = (dd.ops.wq[None, :] * dd.geom.Jq).reshape(-1)
    err = (w * np.abs(rho - ref)).sum() / (w * np.abs(ref)).sum()
    assert err < 1.2e-2, err
    assert rho.min() > 0.4 and st.reduce(T.REDUCE_MIN_RHOE) > 0
    st.close()


@pytest.mark.parametrize("limiter", [SubcellLimiter(), ZhangShuLimiter()], ids=["subcell", "zhangshu"])
def test_1d_leblanc_gpu(limiter):
    """Leblanc tube (density ratio 1e3, pressure ratio 1e9; examples/convergence/leblanc-convergence.jl), N=2, K=100: the
    oracle's state after the first 40 steps, then the GPU run to T = 2/3: positivity and the L1 error against the exact
    solution through calculate_error (the oracle reaches 0.034 summed over the three components)."""
    from p2de_b200.api import State, calculate_error
    from p2de_b200.types import Solver
    param, Ug, Uo, st, orc = run_both(P.leblanc(N=2, K=100, limiter=limiter), 40)
    assert rel(Ug, Uo) < 1e-7
    st.close()
    prm, rd, md, dd, bc, U0 = P.setup(P.leblanc(N=2, K=100, limiter=limiter))
    solver = Solver(param=prm, rd=rd, md=md, discrete_data=dd)
    st = State(solver, bc)
    st.set_state(U0)
    tend, dth = st.ssp33_run(prm.timestepping_param.t0)
    assert abs(tend - prm.timestepping_param.T) < 1e-10 and len(dth) > 1000
    U = st.preallocation.Uq
    assert (U[..., 0] > 0).all() and st.reduce(T.REDUCE_MIN_RHOE) > 0
    err = calculate_error(st, solver, P.leblanc_exact)
    assert err.L1err < 0.06, err
    st.close()
