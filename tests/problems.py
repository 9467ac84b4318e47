"""Problem setups shared by the tests, bench.py and smoke(): the reference's own scenarios
(test/test_smoke.jl, examples/2D/*.jl) plus the synthetic double-Mach-reflection data of
SURVEY.md §8d.  Pure host-side code (numpy); no oracle, no GPU."""
from __future__ import annotations

import math

import numpy as np

from p2de_b200 import (BCData, CompressibleEulerIdealGas, ESLimitedLowOrderPos, GlobalConstant,
                       LaxFriedrichsOnNodalVal, LaxFriedrichsOnProjectedVal, LimitingParameter,
                       LobattoCollocation, NoEntropyProjectionLimiter, Param, PositivityBound,
                       PostprocessingParameter, SubcellLimiter, TimesteppingParameter, ZhangShuLimiter,
                       initialize_data, make_periodic, primitive_to_conservative, sample_initial_condition)


def make_param(N, K, xL, xR, *, limiter=None, rhs=None, T=1.0, CFL=0.5, dt0=1e-2, t0=0.0, gamma=1.4,
               zeta=0.1, eta=0.5, basis=None, dim=2, entropyproj_limiter=None):
    return Param(N=N, K=K, xL=xL, xR=xR,
                 global_constants=GlobalConstant(POSTOL=1e-14, ZEROTOL=5e-16),
                 timestepping_param=TimesteppingParameter(T=T, CFL=CFL, dt0=dt0, t0=t0),
                 limiting_param=LimitingParameter(zeta=zeta, eta=eta),
                 postprocessing_param=PostprocessingParameter(output_interval=10000),
                 equation=CompressibleEulerIdealGas(dim, gamma),
                 rhs=rhs if rhs is not None else ESLimitedLowOrderPos(LaxFriedrichsOnNodalVal(), LaxFriedrichsOnProjectedVal()),
                 approximation_basis=basis if basis is not None else LobattoCollocation(),
                 entropyproj_limiter=entropyproj_limiter if entropyproj_limiter is not None else NoEntropyProjectionLimiter(),
                 rhs_limiter=limiter if limiter is not None else SubcellLimiter(bound=PositivityBound()))


# ---- isentropic vortex: test/test_smoke.jl:6-20 -------------------------------------------
def vortex_exact(eqn, x, y, t):
    gamma = eqn.gamma
    x0, y0, beta = 4.5, 5.0, 8.5
    r2 = (x - x0 - t) ** 2 + (y - y0) ** 2
    u = 1 - beta * np.exp(1 - r2) * (y - y0) / (2 * np.pi)
    v = beta * np.exp(1 - r2) * (x - x0 - t) / (2 * np.pi)
    rho = 1 - (1 / (8 * gamma * np.pi ** 2)) * (gamma - 1) / 2 * (beta * np.exp(1 - r2)) ** 2
    rho = rho ** (1 / (gamma - 1))
    p = rho ** gamma
    return (rho, u, v, p)


def vortex_ic(param, x, y):
    return primitive_to_conservative(param.equation, vortex_exact(param.equation, x, y, param.timestepping_param.t0))


def periodic_bc(param, md):
    return BCData(make_periodic(md).mapP, [], [], [])


def vortex(N=3, K=(5, 5), **kw):
    """test/test_smoke.jl:53-67 (T=2e-2, CFL=1, dt0=1e-2 by default there)."""
    kw.setdefault("T", 2e-2); kw.setdefault("CFL", 1.0); kw.setdefault("dt0", 1e-2)
    return make_param(N, K, (0.0, 0.0), (10.0, 10.0), **kw), vortex_ic, periodic_bc


# ---- Kelvin-Helmholtz: examples/2D/kelvin-helmholtz.jl:9-18 ------------------------------
def kh_ic(param, x, y):
    B = np.tanh(15 * y + 7.5) - np.tanh(15 * y - 7.5)
    return primitive_to_conservative(param.equation, (0.5 + 0.75 * B, 0.5 * (B - 1), 0.1 * np.sin(2 * np.pi * x), 1.0 + 0 * x))


def kelvin_helmholtz(N=3, K=(16, 16), **kw):
    kw.setdefault("T", 10.0); kw.setdefault("CFL", 0.5); kw.setdefault("dt0", 1e-3)
    return make_param(N, K, (-1.0, -1.0), (1.0, 1.0), **kw), kh_ic, periodic_bc


# ---- smooth, plateau-free periodic data (not a reference example): every field varies everywhere and no two
#      stencil nodes carry equal values, so the TVD bounds' min/max and the entropy estimates are decided well away
#      from rounding noise (the KH and vortex data have exactly-flat far fields where they are not)
def wave2d_ic(param, x, y):
    rho = 1.0 + 0.45 * np.sin(2 * np.pi * x + 1.0) * np.cos(2 * np.pi * y + 0.5) + 0.2 * np.cos(4 * np.pi * (x - 0.3 * y) + 0.2)
    u = 0.6 + 0.3 * np.cos(2 * np.pi * y + 0.3)
    v = -0.2 + 0.25 * np.sin(2 * np.pi * x - 0.7)
    p = 1.0 + 0.35 * np.cos(2 * np.pi * (x + y) + 0.9)
    return primitive_to_conservative(param.equation, (rho, u, v, p))


def wave2d(N=3, K=(6, 5), **kw):
    kw.setdefault("T", 1.0); kw.setdefault("CFL", 0.5); kw.setdefault("dt0", 5e-3)
    return make_param(N, K, (0.0, 0.0), (1.0, 1.0), **kw), wave2d_ic, periodic_bc


# ---- Sedov-type blast: examples/2D/sedov.jl:9-20 ------------------------------------------
def sedov(N=3, K=(16, 16), **kw):
    kw.setdefault("T", 1.0); kw.setdefault("CFL", 0.5); kw.setdefault("dt0", 1e-3)
    K1D = K[0]

    def ic(param, x, y):
        g = param.equation.gamma
        r_ini = 4 * 3 / K1D
        r = np.sqrt(x ** 2 + y ** 2)
        p = np.where(r < r_ini, (g - 1) / math.pi / r_ini / r_ini, 1e-5)
        return primitive_to_conservative(param.equation, (1.0 + 0 * x, 0 * x, 0 * x, p))
    return make_param(N, K, (-1.5, -1.5), (1.5, 1.5), **kw), ic, periodic_bc


# ---- synthetic double Mach reflection data (SURVEY.md §8d "S-DMR") -------------------------
DMR_POST = (8.0, 8.25 * math.cos(math.pi / 6), -8.25 * math.sin(math.pi / 6), 116.5)
DMR_PRE = (1.4, 0.0, 0.0, 1.0)


def dmr_ic(param, x, y):
    post = x < 1.0 / 6.0 + y / math.sqrt(3.0)
    prim = tuple(np.where(post, a, b) for a, b in zip(DMR_POST, DMR_PRE))
    return primitive_to_conservative(param.equation, prim)


def dmr_bc(param, md):
    """Left: Dirichlet inflow at the post-shock state (mapI/Ival); right, bottom, top: copy-out
    (mapO), through the reference's BC surface (src/common/types/StateParam.jl:1-7)."""
    eq = param.equation
    Nfp = md.mapP.shape[1]
    n = Nfp // 4
    Kx, Ky = md.Kxy
    k = np.arange(md.K)
    ix, iy = k % Kx, k // Kx
    idx = md.mapM                                    # 1-based linear indices [K, Nfp]
    left = idx[ix == 0, 0:n].reshape(-1)
    right = idx[ix == Kx - 1, n:2 * n].reshape(-1)
    bottom = idx[iy == 0, 2 * n:3 * n].reshape(-1)
    top = idx[iy == Ky - 1, 3 * n:4 * n].reshape(-1)
    Ival = np.tile(np.array(primitive_to_conservative(eq, DMR_POST)), (len(left), 1))
    return BCData(md.mapP, left, np.concatenate([right, bottom, top]), Ival)


def dmr(N=3, K=(64, 16), **kw):
    kw.setdefault("T", 0.2); kw.setdefault("CFL", 0.5); kw.setdefault("dt0", 1e-3)
    return make_param(N, K, (0.0, 0.0), (4.0, 1.0), **kw), dmr_ic, dmr_bc


# ---- 1D shock tubes (BASELINE.json configs 1-2; the reference ships no Sod script, SURVEY.md §0.1:
#      standard data, inflow left via mapI=[1], outflow right via mapO=[2K] as in
#      examples/convergence/leblanc-convergence.jl:45-52) ------------------------------------
def sod(N=3, K=200, **kw):
    kw.setdefault("T", 0.2); kw.setdefault("CFL", 0.5); kw.setdefault("dt0", 1e-3)

    def ic(param, x):
        left = x < 0.5
        return primitive_to_conservative(param.equation, (np.where(left, 1.0, 0.125), 0 * x, np.where(left, 1.0, 0.1)))

    def bc(param, md):
        Ival = np.array([primitive_to_conservative(param.equation, (1.0, 0.0, 1.0))])
        return BCData(md.mapP, [1], [2 * md.K], Ival)
    return make_param(N, K, 0.0, 1.0, dim=1, **kw), ic, bc


def shu_osher(N=3, K=64, **kw):
    """examples/1D/shu-osher.jl:9-30: Mach-3 shock into a density sine wave on [-5, 5]."""
    kw.setdefault("T", 1.8); kw.setdefault("CFL", 0.5); kw.setdefault("dt0", 1e-3)
    post = (3.857143, 2.629369, 10.3333)

    def ic(param, x):
        left = x < -4.0
        return primitive_to_conservative(param.equation, (np.where(left, post[0], 1 + 0.2 * np.sin(5 * x)),
                                                         np.where(left, post[1], 0.0), np.where(left, post[2], 1.0)))

    def bc(param, md):
        Ival = np.array([primitive_to_conservative(param.equation, post)])
        return BCData(md.mapP, [1], [2 * md.K], Ival)
    return make_param(N, K, -5.0, 5.0, dim=1, **kw), ic, bc


# ---- Leblanc shock tube: examples/convergence/leblanc-convergence.jl:9-62 (gamma = 5/3, t0 = 0.01, T = 2/3) -------------
def leblanc_exact(eqn, x, t):
    """exact_sol of the reference script, vectorised: primitives (rho, u, p) at time t."""
    g = eqn.gamma
    x = np.asarray(x, dtype=np.float64)
    rhoL, rhoR, pL, pR = 1.0, 1e-3, (g - 1) * 1e-1, (g - 1) * 1e-10
    if t == 0:
        left = x < 0.33
        return np.where(left, rhoL, rhoR), 0 * x, np.where(left, pL, pR)
    xi = (x - 0.33) / t
    rhoLs, rhoRs = 5.4079335349316249e-2, 3.9999980604299963e-3
    vs, ps = 0.62183867139173454, 0.51557792765096996e-3
    l1, l3 = 0.49578489518897934, 0.82911836253346982
    fan = np.clip(0.75 - 0.75 * xi, 1e-300, None)
    rho = np.select([xi <= -1 / 3, xi <= l1, xi <= vs, xi <= l3], [rhoL, fan ** 3, rhoLs, rhoRs], rhoR)
    u = np.select([xi <= -1 / 3, xi <= l1, xi <= l3], [0.0, 0.75 * (1 / 3 + xi), vs], 0.0)
    p = np.select([xi <= -1 / 3, xi <= l1, xi <= l3], [pL, fan ** 5 / 15, ps], pR)
    return rho, u, p


def leblanc(N=2, K=100, **kw):
    kw.setdefault("T", 2.0 / 3.0); kw.setdefault("CFL", 0.5); kw.setdefault("dt0", 1e-3); kw.setdefault("t0", 0.01)
    kw.setdefault("gamma", 5.0 / 3.0); kw.setdefault("eta", 0.1)

    def ic(param, x):
        return primitive_to_conservative(param.equation, leblanc_exact(param.equation, x, param.timestepping_param.t0))

    def bc(param, md):
        Ival = np.array([primitive_to_conservative(param.equation, (1.0, 0.0, (param.equation.gamma - 1) * 1e-1))])
        return BCData(md.mapP, [1], [2 * md.K], Ival)
    return make_param(N, K, 0.0, 1.0, dim=1, **kw), ic, bc


def density_wave_1d(N=3, K=16, **kw):
    kw.setdefault("T", 1.0); kw.setdefault("CFL", 0.5); kw.setdefault("dt0", 1e-2)

    def ic(param, x):
        return primitive_to_conservative(param.equation, (1 + 0.5 * np.sin(2 * np.pi * x), 0.7 + 0 * x, 1.0 + 0 * x))

    def bc(param, md):
        return BCData(make_periodic(md, (True, False)).mapP, [], [], [])
    return make_param(N, K, 0.0, 1.0, dim=1, **kw), ic, bc


def setup(problem):
    """(param, ic, bc_callback) -> (param, rd, md, discrete_data, bcdata, U0)."""
    param, ic, bcf = problem
    rd, md, dd = initialize_data(param)
    bc = bcf(param, md)
    U0 = sample_initial_condition(param, md, ic)
    return param, rd, md, dd, bc, U0
