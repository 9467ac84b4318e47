"""Known-answer tests that pin the oracle's pointwise physics to the reference's formulas
(src/math/compressible_Navier_Stokes.jl, src/dg/limiter/limiter_utils.jl)."""
import ctypes as C
import math

import numpy as np

from oracle.oracle import lib

GAMMA = 1.4


def arr(*v):
    return np.array(v, dtype=np.float64)


def cons(rho, u, v, p):
    return arr(rho, rho * u, rho * v, p / (GAMMA - 1) + 0.5 * rho * (u * u + v * v))


def test_logmean_series_constants_and_branch():
    """compressible_Navier_Stokes.jl:307-321: |f| < 1e-4 -> aavg*(1 + v*(-0.2 - v*(0.0512 - v*0.026038857142857)))."""
    L = lib()
    aL, aR = 1.0, 1.0 + 5e-5
    da, aavg = aR - aL, 0.5 * (aR + aL)
    v = (da / aavg) ** 2
    assert L.oracle_logmean(aL, aR) == aavg * (1 + v * (-0.2 - v * (0.0512 - v * 0.026038857142857)))
    aR = 1.3
    assert L.oracle_logmean(aL, aR) == -(aR - aL) / (math.log(aL) - math.log(aR))
    assert L.oracle_logmean(2.5, 2.5) == 2.5
    assert L.oracle_logmean(0.7, 1.9) == L.oracle_logmean(1.9, 0.7)      # bitwise symmetric


def test_entropy_variable_roundtrip_and_values():
    """v_ufun / u_vfun (:134-163): u(v(U)) = U, and v4 = -rho (gamma-1)/p."""
    L = lib()
    U = cons(1.3, 0.4, -0.7, 2.1)
    V, W = np.empty(4), np.empty(4)
    L.oracle_v_ufun_2d(GAMMA, U.ctypes.data, V.ctypes.data)
    assert abs(V[3] + 1.3 * (GAMMA - 1) / 2.1) < 1e-15
    s = math.log(2.1 / 1.3 ** GAMMA)
    assert abs(V[0] - ((GAMMA + 1 - s) - (GAMMA - 1) * U[3] / 2.1)) < 1e-14
    L.oracle_u_vfun_2d(GAMMA, V.ctypes.data, W.ctypes.data)
    assert np.abs(W - U).max() < 1e-14


def test_two_point_flux_consistency_and_symmetry():
    """fS (:220-249): fS(U, U) = f(U); fS(UL, UR) = fS(UR, UL)."""
    L = lib()
    U, W = cons(1.1, 0.3, 0.2, 0.9), cons(0.4, -1.0, 0.5, 2.0)
    F, G, H = np.empty(8), np.empty(8), np.empty(8)
    L.oracle_fS_2d(GAMMA, U.ctypes.data, U.ctypes.data, F.ctypes.data)
    L.oracle_fluxes_2d(GAMMA, U.ctypes.data, G.ctypes.data)
    assert np.abs(F - G).max() < 1e-14
    L.oracle_fS_2d(GAMMA, U.ctypes.data, W.ctypes.data, F.ctypes.data)
    L.oracle_fS_2d(GAMMA, W.ctypes.data, U.ctypes.data, H.ctypes.data)
    assert np.array_equal(F, H)
    # hand evaluation of the x-mass flux: logmean(rho) * avg(u)
    rl = -(0.4 - 1.1) / (math.log(1.1) - math.log(0.4))
    assert abs(F[0] - rl * 0.5 * (0.3 - 1.0)) < 1e-15


def lp(U, P, Lrho, Lrhoe, Urho=math.inf, Urhoe=math.inf, ZT=5e-16):
    return lib().oracle_limiting_param_2d(ZT, U.ctypes.data, P.ctypes.data, Lrho, Lrhoe, Urho, Urhoe)


def test_limiting_param_density_and_energy_bounds():
    """limiting_param_bound_rho_rhoe / rhoe_quadratic_solve (limiter_utils.jl:26-76)."""
    U = cons(1.0, 0.0, 0.0, 1.0)
    rhoe = U[3]
    # inside the bounds: l = 1
    assert lp(U, arr(0.1, 0.0, 0.0, 0.1), 0.1, 0.1 * rhoe) == 1.0
    # density bound: rho + l*P = Lrho  ->  l = (Lrho - rho)/P
    assert lp(U, arr(-2.0, 0.0, 0.0, 0.0), 0.1, -1.0) == (0.1 - 1.0) / -2.0
    # energy bound with P only in E: rho*(E + l PE) - rho*L = 0 -> l = (L - E)/PE  (a = 0 -> roots +-Inf/NaN -> 1?)
    # a = P1*P4 - |Pm|^2/2 = 0 here: the reference falls through to l = 1
    assert lp(U, arr(0.0, 0.0, 0.0, -5.0), 0.1, 0.1 * rhoe) == 1.0
    # genuine quadratic: momentum perturbation drives rho*e down
    P = arr(0.0, 3.0, 0.0, 0.0)
    l = lp(U, P, 0.1, 0.1 * rhoe)
    assert 0 < l < 1
    Ul = U + l * P
    assert abs((Ul[3] - 0.5 * (Ul[1] ** 2 + Ul[2] ** 2) / Ul[0]) - 0.1 * rhoe) < 1e-13
    # Lrhoe == Inf returns 1 immediately
    assert lp(U, P, 0.1, math.inf) == 1.0
