"""Oracle checks for the remaining subcell bounds: TVD bounds (subcell.jl:77-141, rho_bound :352-361)
and cell-entropy bounds (enforce_ES_subcell!, subcell.jl:458-716).  The reference pins no numbers for them
(test/test_smoke.jl:50-51 only runs two of them), so the restatement is validated by the properties the bounds exist
to enforce."""
import numpy as np
import pytest

import problems as P
from oracle.oracle import Oracle
from p2de_b200 import (PositivityAndCellEntropyBound, PositivityAndRelaxedCellEntropyBound, PositivityBound,
                       SubcellLimiter, TVDAndCellEntropyBound, TVDAndMinEntropyBound, TVDAndRelaxedCellEntropyBound,
                       TVDBound, primitive_to_conservative)

CELL = [PositivityAndCellEntropyBound(), PositivityAndRelaxedCellEntropyBound(beta=0.5), TVDAndCellEntropyBound(),
        TVDAndRelaxedCellEntropyBound(beta=0.5)]
TVD = [TVDBound(), TVDAndCellEntropyBound(), TVDAndMinEntropyBound()]


def one_rhs(problem, dt=None, nstage=1):
    param, rd, md, dd, bc, U0 = P.setup(problem)
    orc = Oracle(param, dd, bc, threads=2)
    orc.set_state(U0)
    tp = param.timestepping_param
    dt = tp.CFL * tp.dt0 if dt is None else dt
    orc.rhs(tp.t0, dt, nstage)
    return param, dd, orc, U0, dt


def interior_blocks(L, N1D):
    """L_local[k, d, :] -> the interior subcell faces in the reference's dvdf index order (subcell.jl:625-631)."""
    Nq = N1D * N1D
    Lx = L[:, 0, :].reshape(-1, N1D, N1D + 1)[:, :, 1:N1D].reshape(-1, Nq - N1D)     # (si-1) + sj (N1D-1)
    Ly = L[:, 1, :].reshape(-1, N1D + 1, N1D)[:, 1:N1D, :].reshape(-1, Nq - N1D)     # si + (sj-1) N1D
    return Lx, Ly


@pytest.mark.parametrize("bound", CELL, ids=lambda b: type(b).__name__)
@pytest.mark.parametrize("problem", ["wave", "kh", "vortex"])
def test_cell_entropy_inequality_holds_after_enforcement(bound, problem):
    """sum_faces l (v_{s-1} - v_s).(f_H - f_L) <= rhs_es + tol per element and direction (subcell.jl:632-650), and the
    coefficients only ever decrease with respect to the same bound without the entropy part."""
    fac = {"wave": lambda lim: P.wave2d(N=3, limiter=lim), "kh": lambda lim: P.kelvin_helmholtz(N=3, K=(8, 8), limiter=lim),
           "vortex": lambda lim: P.vortex(N=2, K=(5, 5), limiter=lim)}[problem]
    param, dd, orc, U0, dt = one_rhs(fac(SubcellLimiter(bound=bound)))
    N1D = param.N + 1
    Nq = N1D * N1D
    L = orc.field("L_local")[0]
    dvx = orc.field("dvdf_x").reshape(-1, Nq)[:, :Nq - N1D]
    dvy = orc.field("dvdf_y").reshape(-1, Nq)[:, :Nq - N1D]
    sB = orc.field("sum_Bpsi").reshape(-1, 2)
    sL = orc.field("sum_dvfbarL").reshape(-1, 2)
    eps = orc.field("smooth_factor").reshape(3, -1)[0]
    relaxed = "Relaxed" in type(bound).__name__
    assert (eps == 0).all() if not relaxed else ((eps >= 0).all() and (eps <= 1).all())
    Lx, Ly = interior_blocks(L, N1D)
    for d, (dv, Ld) in enumerate(((dvx, Lx), (dvy, Ly))):
        rhs = (1 - 0.5 * eps if relaxed else 1.0) * (sB[:, d] - sL[:, d])
        tol = np.maximum(0.0, sL[:, d] - sB[:, d])
        lhs = (Ld * dv).sum(1)
        scale = np.abs(Ld * dv).sum(1) + np.abs(rhs) + 1e-300
        # elements where the greedy loop stopped at dvdf < ZEROTOL (subcell.jl:682-684) may keep a violation
        stuck = (dv < param.global_constants.ZEROTOL).all(1)
        assert ((lhs - rhs - tol) <= 1e-12 * scale)[~stuck].all()
    base = TVDBound() if type(bound).__name__.startswith("TVD") else PositivityBound()
    _, _, orc0, _, _ = one_rhs(fac(SubcellLimiter(bound=base)))
    L0 = orc0.field("L_local")[0]
    assert (L <= L0 + 1e-15).all() and (L >= 0).all()
    assert (L < L0).any()                   # ... and the bound does bite on this data
    # the low- and high-order parts do not depend on the bound
    assert np.array_equal(orc.field("rhsL"), orc0.field("rhsL")) and np.array_equal(orc.field("rhsH"), orc0.field("rhsH"))


def test_cell_entropy_free_stream_untouched():
    param, ic, bc = P.vortex(N=3, K=(4, 3), limiter=SubcellLimiter(bound=PositivityAndCellEntropyBound()))
    const = lambda prm, x, y: primitive_to_conservative(prm.equation, (1.2 + 0 * x, 0.3 + 0 * x, -0.4 + 0 * x, 0.9 + 0 * x))
    param, dd, orc, U0, dt = one_rhs((param, const, bc))
    assert (orc.field("L_local")[0] == 1.0).all()
    assert np.abs(orc.field("rhsU")).max() < 1e-12


@pytest.mark.parametrize("bound", TVD, ids=lambda b: type(b).__name__)
@pytest.mark.parametrize("N", [1, 3])
def test_tvd_bounds_hold_for_the_limited_update(bound, N):
    """rho(Uq + dt rhsU) stays inside [min, max] of rho(Uq + dt rhsL) over the low-order stencil: the limited update is
    the average of 2*Nd sub-updates uL + l P, each of which the limiter keeps inside the bounds."""
    param, dd, orc, U0, dt = one_rhs(P.wave2d(N=N, limiter=SubcellLimiter(bound=bound)), dt=4e-3)
    rho_new = (U0 + dt * orc.field("rhsU"))[:, :, 0]
    lb = orc.field("lbound_rho").reshape(rho_new.shape)
    ub = orc.field("ubound_rho").reshape(rho_new.shape)
    rhoL = (U0 + dt * orc.field("rhsL"))[:, :, 0]
    assert (lb <= rhoL).all() and (rhoL <= ub).all()
    assert (rho_new >= lb - 1e-13).all() and (rho_new <= ub + 1e-13).all()
    # without the TVD part the same update leaves the bounds somewhere (the test would be vacuous otherwise)
    _, _, orc0, _, _ = one_rhs(P.wave2d(N=N, limiter=SubcellLimiter(bound=PositivityBound())), dt=4e-3)
    rho0 = (U0 + dt * orc0.field("rhsU"))[:, :, 0]
    assert ((rho0 < lb - 1e-10) | (rho0 > ub + 1e-10)).any()
    assert (orc.field("L_local")[0] < 1).any()


@pytest.mark.parametrize("bound", [TVDBound(), PositivityAndCellEntropyBound(), TVDAndRelaxedCellEntropyBound(beta=0.5)],
                         ids=lambda b: type(b).__name__)
def test_new_bounds_keep_conservation_and_symmetric_interfaces(bound):
    param, dd, orc, U0, dt = one_rhs(P.wave2d(N=2, K=(5, 4), limiter=SubcellLimiter(bound=bound)))
    wJ = dd.ops.wq[None, :, None] * dd.geom.Jq[:, :, None]
    rU = orc.field("rhsU")
    assert np.abs((wJ * rU).sum((0, 1))).max() < 1e-13 * (wJ * np.abs(rU)).sum((0, 1)).max()
    N1D, (Kx, Ky) = param.N + 1, param.K
    L = orc.field("L_local")[0].reshape(Ky, Kx, 2, -1)
    Lx = L[:, :, 0, :].reshape(Ky, Kx, N1D, N1D + 1)
    Ly = L[:, :, 1, :].reshape(Ky, Kx, N1D + 1, N1D)
    assert np.array_equal(Lx[:, :, :, N1D], np.roll(Lx[:, :, :, 0], -1, axis=1))
    assert np.array_equal(Ly[:, :, N1D, :], np.roll(Ly[:, :, 0, :], -1, axis=0))


# ---- 1D (subcell.jl:37-53,93-118,466-500,568-610: the Dim1 methods of the same bounds; no GPU kernel yet) ------------------
@pytest.mark.parametrize("bound", [TVDBound(), TVDAndMinEntropyBound()], ids=lambda b: type(b).__name__)
@pytest.mark.parametrize("problem", ["shu-osher", "wave"])
def test_1d_tvd_bounds_hold_for_the_limited_update(bound, problem):
    prob = P.shu_osher(N=3, K=64, limiter=SubcellLimiter(bound=bound)) if problem == "shu-osher" else \
        P.density_wave_1d(N=3, K=16, limiter=SubcellLimiter(bound=bound))
    param, dd, orc, U0, dt = one_rhs(prob)
    rho_new = (U0 + dt * orc.field("rhsU"))[:, :, 0]
    lb = orc.field("lbound_rho").reshape(rho_new.shape)
    ub = orc.field("ubound_rho").reshape(rho_new.shape)
    assert (rho_new >= lb - 1e-13).all() and (rho_new <= ub + 1e-13).all()
    assert (orc.field("L_local")[0][:, 0, :param.N + 2] < 1).any()


@pytest.mark.parametrize("bound", [PositivityAndCellEntropyBound(), PositivityAndRelaxedCellEntropyBound(beta=0.5)],
                         ids=lambda b: type(b).__name__)
def test_1d_cell_entropy_inequality_holds_after_enforcement(bound):
    param, dd, orc, U0, dt = one_rhs(P.density_wave_1d(N=3, K=16, limiter=SubcellLimiter(bound=bound)))
    N1D = param.N + 1
    L = orc.field("L_local")[0][:, 0, :N1D + 1]
    dv = orc.field("dvdf_x").reshape(-1, N1D)[:, :N1D - 1]
    sB = orc.field("sum_Bpsi").reshape(-1)
    sL = orc.field("sum_dvfbarL").reshape(-1)
    eps = orc.field("smooth_factor").reshape(3, -1)[0]
    rhs = (1 - 0.5 * eps if "Relaxed" in type(bound).__name__ else 1.0) * (sB - sL)
    tol = np.maximum(0.0, sL - sB)
    lhs = (L[:, 1:N1D] * dv).sum(1)
    assert (lhs - rhs - tol <= 1e-12 * (np.abs(L[:, 1:N1D] * dv).sum(1) + np.abs(rhs) + 1e-300)).all()
    assert (L < 1).any() and (L >= 0).all()


# ---- x <-> y symmetry: the y-direction code of every bound against its x-direction code ------------------------------------
def _transpose_problem(problem):
    """The same flow with the roles of x and y exchanged: K = (Ky, Kx), data(x, y) = swap_uv(data0(y, x))."""
    param, ic, bc = problem
    import dataclasses
    pT = dataclasses.replace(param, K=(param.K[1], param.K[0]), xL=(param.xL[1], param.xL[0]), xR=(param.xR[1], param.xR[0]))

    def icT(prm, x, y):
        U = ic(param, y, x)
        return (U[0], U[2], U[1], U[3])
    return pT, icT, bc


@pytest.mark.parametrize("bound", [PositivityBound(), TVDBound(), PositivityAndCellEntropyBound(),
                                   TVDAndRelaxedCellEntropyBound(beta=0.5), TVDAndMinEntropyBound()], ids=lambda b: type(b).__name__)
def test_bounds_are_symmetric_under_exchange_of_x_and_y(bound):
    prob = P.wave2d(N=3, K=(5, 4), limiter=SubcellLimiter(bound=bound))
    param, dd, orc, U0, dt = one_rhs(prob)
    paramT, ddT, orcT, U0T, _ = one_rhs(_transpose_problem(prob))
    n = param.N + 1
    Kx, Ky = param.K

    def to_T(F):   # [K, Nq, 4] of the original -> layout of the transposed problem
        G = F.reshape(Ky, Kx, n, n, 4).transpose(1, 0, 3, 2, 4)        # element (iy, ix) -> (ix, iy); node (j, i) -> (i, j)
        return G[..., [0, 2, 1, 3]].reshape(Kx * Ky, n * n, 4)
    assert np.abs(to_T(U0) - U0T).max() < 1e-14
    for f in ("rhsL", "rhsH", "rhsU"):
        a, b = to_T(orc.field(f)), orcT.field(f)
        assert np.abs(a - b).max() < 1e-11 * np.abs(b).max(), f
    L, LT = orc.field("L_local")[0].reshape(Ky, Kx, 2, -1), orcT.field("L_local")[0].reshape(Kx, Ky, 2, -1)
    Lx = L[:, :, 0, :].reshape(Ky, Kx, n, n + 1)          # [iy, ix, sj, si]
    Ly = L[:, :, 1, :].reshape(Ky, Kx, n + 1, n)          # [iy, ix, sj, si]
    LTx = LT[:, :, 0, :].reshape(Kx, Ky, n, n + 1)
    LTy = LT[:, :, 1, :].reshape(Kx, Ky, n + 1, n)
    # the original's x faces (si, sj) are the transposed problem's y faces (si' = sj, sj' = si) and vice versa; faces whose
    # f_bar_H - f_bar_L is rounding noise are excluded (TVD coefficients there are 0 or 1 by the sign of the noise)
    def sig(o, ax, shape):
        fH = o.field(f"f_bar_H_{ax}").reshape(-1, n * n + n, 4); fL = o.field(f"f_bar_L_{ax}").reshape(-1, n * n + n, 4)
        return (np.abs(fH - fL).max(-1) > 1e-10 * np.abs(fH).max()).reshape(shape)
    sx = sig(orc, "x", (Ky, Kx, n, n + 1))
    sy = sig(orc, "y", (Ky, Kx, n + 1, n))
    dx = (Lx.transpose(1, 0, 3, 2) - LTy) * sx.transpose(1, 0, 3, 2)
    dy = (Ly.transpose(1, 0, 3, 2) - LTx) * sy.transpose(1, 0, 3, 2)
    assert np.abs(dx).max() < 1e-9 and np.abs(dy).max() < 1e-9
