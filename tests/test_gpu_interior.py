"""GPU parity on meshes wide enough for the kernel instantiations the benchmark times.

A CTA of the FAST stage kernel runs its compile-time INTERIOR version only when its batch of 16 (N=4: 12) elements lies
strictly inside a structured mesh (csrc/stage_fast.cuh: fast_batch), i.e. with at least 3 batches per element row and 3
rows; stage 2 of the default schedule additionally runs the deferred-combine instantiation.  The meshes of
test_gpu_parity.py are too small for either, so the same comparisons against the CPU oracle are repeated here on wide
meshes, and p2de_debug_counters proves which instantiations ran.  Reference: flux_differencing.jl:164-211,
subcell.jl:248-349,418-456, SSPRK33.jl:28-40.
"""
import numpy as np
import pytest

import problems as P
from p2de_b200 import TimeParam
from test_gpu_parity import make_pair, rel

pytestmark = pytest.mark.gpu

CASES = {
    "dmr-N3-64x8": lambda: P.dmr(N=3, K=(64, 8)),            # 2D grid of CTAs (Kx a multiple of the batch size)
    "dmr-N3-80x6": lambda: P.dmr(N=3, K=(80, 6)),
    "dmr-N3-72x6": lambda: P.dmr(N=3, K=(72, 6)),            # Kx not a multiple of 16: the 1D-grid branch of fast_batch
    "kh-N3-64x6": lambda: P.kelvin_helmholtz(N=3, K=(64, 6)),
    "kh-N4-60x6": lambda: P.kelvin_helmholtz(N=4, K=(60, 6)),  # 12 elements per CTA
    "kh-N4-66x5": lambda: P.kelvin_helmholtz(N=4, K=(66, 5)),
    "wave-N3-64x5": lambda: P.wave2d(N=3, K=(64, 5)),
    "vortex-N1-64x4": lambda: P.vortex(N=1, K=(64, 4), T=10.0),   # (T far away: the smoke test's T = 2e-2 is reached after two steps)
    "vortex-N2-64x4": lambda: P.vortex(N=2, K=(64, 4), T=10.0),
    "sedov-N3-64x6": lambda: P.sedov(N=3, K=(64, 6)),
}
DT = {"dmr": 5e-4, "sedov": 2e-2}


def formulation_sensitivity(problem, nstage, dt):
    """How far the reference formulation itself moves under a legal re-association of its arithmetic: the oracle compiled
    with FMA contraction against the oracle compiled without (same source, oracle/Makefile).  On fine meshes logmean's
    log branch (-da / (logL - logR) with |da/a| ~ 1e-4..1e-3) has a relative rounding error ~1e-16 / |da/a| that the
    near-cancelling sum of S_ij F_ij amplifies, so 1e-12 is below what ANY two builds of the reference agree to there."""
    from oracle.oracle import Oracle
    param, rd, md, dd, bc, U0 = P.setup(problem)
    out = []
    # "series": the reference's -da / (log aL - log aR) against the same mean evaluated without the cancellation (the CUDA
    # path uses that series where no pair of a line differs by more than 9 %, csrc/stage_fast.cuh: fS_rot_smooth)
    for variant in ("ref", "fma", "series"):
        o = Oracle(param, dd, bc, variant=variant)
        o.set_state(U0)
        o.rhs(param.timestepping_param.t0, dt, nstage)
        out.append({k: o.field(k) for k in ("rhsU", "rhsH", "rhsL")})
    return {k: max(rel(out[1][k], out[0][k]), rel(out[2][k], out[0][k])) for k in out[0]}


def _dt(name, param):
    tp = param.timestepping_param
    return DT.get(name.split("-")[0], tp.CFL * tp.dt0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_interior_rhs_per_stage(name):
    """One rhs! per stage index, same tolerances as test_gpu_parity.check_rhs, INTERIOR CTAs counted."""
    from p2de_b200.api import rhs
    for nstage in (1, 2, 3):
        param, solver, st, orc, U0 = make_pair(CASES[name]())
        dt = _dt(name, param)
        st.debug_counters(True)
        tp = param.timestepping_param
        dt_o = orc.rhs(tp.t0, dt, nstage)
        dt_g = rhs(st, solver, None, TimeParam(t=tp.t0, dt=dt, nstage=nstage))
        cnt = st.debug_counters(False)
        assert cnt["cta_interior"] > 0 and cnt["cta_general"] > 0, cnt
        assert abs(dt_g - dt_o) <= 1e-13 * abs(dt_o), (dt_g, dt_o)
        pre = st.preallocation
        # 1e-12 relative, or four times the formulation's own sensitivity where that is larger (fine meshes, see above)
        sens = formulation_sensitivity(CASES[name](), nstage, dt)
        for f in ("rhsU", "rhsL", "rhsH"):
            assert rel(getattr(pre, f), orc.field(f)) < max(1e-12, 4 * sens[f]), (f, sens)
        Lg, Lo = pre.L_local[nstage - 1], orc.field("L_local")[nstage - 1]
        assert np.abs(Lg - Lo).max() < 1e-12
        assert np.array_equal(Lg == 1.0, Lo == 1.0)
        if name.startswith(("dmr", "sedov")):
            assert (Lo < 1.0).any()          # the limiter bites on these
        st.close()


STEP_TOL = {"dmr": 1e-7, "sedov": 1e-7}


@pytest.mark.parametrize("name", sorted(CASES))
def test_interior_ssp33_steps_default_schedule(name):
    """12 SSP33! steps through the default 4-launch schedule (stage 2 = the deferred-combine kernel): states against
    the oracle, identical signs of rho and rho e."""
    param, solver, st, orc, U0 = make_pair(CASES[name](), keep_diagnostics=False)
    st.debug_counters(True)
    t_o = t_g = param.timestepping_param.t0
    l0 = st.kernel_launch_count()
    nsteps = 12
    for _ in range(nsteps):
        dto = orc.ssp33_step(t_o); t_o += dto
        dtg = st.ssp33_step(t_g); t_g += dtg
        assert abs(dtg - dto) <= 1e-10 * dto
    cnt = st.debug_counters(False)
    assert st.kernel_launch_count() - l0 == 4 * nsteps          # set_dt + three stage kernels, nothing else
    assert cnt["cta_interior"] > 0 and cnt["cta_defer"] > 0, cnt
    ctas = cnt["cta_interior"] + cnt["cta_general"]
    assert cnt["cta_defer"] * 3 == ctas                           # exactly one of the three stages defers
    Ug, Uo = st.preallocation.Uq, orc.get_state()
    assert rel(Ug, Uo) < STEP_TOL.get(name.split("-")[0], 1e-9)
    rhoe = lambda U: U[..., 3] - 0.5 * (U[..., 1] ** 2 + U[..., 2] ** 2) / U[..., 0]
    assert np.array_equal(np.sign(Ug[..., 0]), np.sign(Uo[..., 0])) and (Ug[..., 0] > 0).all()
    assert np.array_equal(np.sign(rhoe(Ug)), np.sign(rhoe(Uo))) and (rhoe(Ug) > 0).all()
    st.close()


def test_interior_step_keeps_all_three_stage_coefficients():
    """SSP33! leaves L_local[:, :, :, 1:3] of the last step behind (SSPRK33.jl:31-39, State.jl:21); with keep_diagnostics
    the default schedule writes each stage's coefficients into its own slot without extra launches."""
    param, solver, st, orc, U0 = make_pair(P.dmr(N=3, K=(64, 8)), keep_diagnostics=True)
    t = param.timestepping_param.t0
    l0 = st.kernel_launch_count()
    for _ in range(3):
        dto = orc.ssp33_step(t)
        st.ssp33_step(t)
        t += dto
    assert st.kernel_launch_count() - l0 == 3 * 4
    Lg, Lo = st.preallocation.L_local, orc.field("L_local")
    for s in range(3):
        # state differences of ~1e-10 after three shock steps move active coefficients a little; the {l == 1} sets agree
        assert np.abs(Lg[s] - Lo[s]).max() < 1e-6, s
        assert (np.asarray(Lg[s] == 1.0) != np.asarray(Lo[s] == 1.0)).mean() < 1e-3
    assert (Lo < 1.0).any()
    st.close()


def test_tiled_developed_shock_state_steps():
    """bench.py's tiled "developed" construction at test size: the double-Mach-reflection data advanced on a small mesh, tiled 2x2
    onto a mesh of twice the extent (strong discontinuities at the tile seams, inflow / outflow boundary conditions), then
    30 SSP33! steps against the oracle: the limiter's exact evaluation and the non-quiet flux path run on most lines."""
    from oracle.oracle import Oracle
    param, rd, md, dd, bc, U0 = P.setup(P.dmr(N=3, K=(64, 16), T=1e9))
    o = Oracle(param, dd, bc)
    o.set_state(U0)
    t = 0.0
    for _ in range(190):
        t += o.ssp33_step(t)
    Us = o.get_state().reshape(16, 64, 16, 4)
    big = np.ascontiguousarray(np.tile(Us, (2, 2, 1, 1)).reshape(-1, 16, 4))
    param, solver, st, orc, _ = make_pair(P.dmr(N=3, K=(128, 32), T=1e9), keep_diagnostics=False)
    st.set_state(big); orc.set_state(big)
    st.debug_counters(True)
    t_o = t_g = 0.0
    for i in range(30):
        dto = orc.ssp33_step(t_o); t_o += dto
        dtg = st.ssp33_step(t_g); t_g += dtg
        assert abs(dtg - dto) <= 1e-9 * dto, (i, dtg, dto)
    cnt = st.debug_counters(False)
    Ug, Uo = st.preallocation.Uq, orc.get_state()
    assert np.isfinite(Ug).all() and (Ug[..., 0] > 0).all()
    assert rel(Ug, Uo) < 1e-6
    assert cnt["lines_not_easy"] > 0 and cnt["limiter_slow"] > 0 and cnt["elem_logs"] > 0
    st.close()
