"""A SECOND, independently written restatement of one `rhs!` of the reference (2D compressible Euler), used only to cross-check
the C++ oracle (tests/test_oracle_crosscheck.py).  It shares no code and no structure with oracle/p2de_oracle.cpp:

* dense operators and plain numpy, vectorised over elements: the flux-differencing volume term is the full Nh x Nh Hadamard sum
  `QF1_i = sum_j Sxyh_db[i, j] * fS(u_i, u_j)` (flux_differencing.jl:164-211 visits the non-zeros of the same matrix and adds the
  skew-symmetric partner), the assembly is the dense `M^-1 Vh^T QF1 + M^-1 Vf^T BF` (flux_differencing.jl:274-361), face data are
  gathered through the caller's linear `mapP`, boundary conditions through `mapI / mapO / Ival`;
* no line structure, no tensor-product knowledge, no special case for Lobatto nodes.

Covered (2D: `dense_rhs`, `dense_limited_rhs`, `dense_ssp33_step`; 1D: `dense_limited_rhs_1d`):
* entropy projection (rhs.jl:59-133), with `NodewiseScaledExtrapolation` on Gauss nodes (`dense_theta`: filter.jl:6-130, the projection
  with theta, the limited face matrix of flux_differencing.jl:288-319);
* `rhs_low_graph_visc!` including the CFL dt and `find_alpha` (low_order_graph_viscosity.jl:4-327);
* `rhs_fluxdiff!` with both volume fluxes and both surface fluxes (flux_differencing.jl:4-361);
* `apply_rhs_limiter!` (limiter.jl:8-56): Zhang-Shu, the subcell limiter with all ten bounds of Solver.jl:47-63 and Hennemann shock
  capturing, as whole-array operations with the interface symmetrisation and the low-order stencils through `mapP`; the cell-entropy
  bounds element by element (`es_volume`; on Gauss nodes also the interface part, `es_interface`, in element order: the reference's
  own result depends on its thread interleaving there, oracle deviation D5);
* the `SSP33!` loop (SSPRK33.jl:28-40).

Test infrastructure; nothing under p2de_b200/ imports it."""
import math

import numpy as np

from p2de_b200 import types as T

_mlog = np.frompyfunc(math.log, 1, 1)
_mexp = np.frompyfunc(math.exp, 1, 1)
_mpow = np.frompyfunc(math.pow, 2, 1)


def LOG(x):
    """libm's log, element by element.  numpy's vectorised log differs from libm's in the last bit for some arguments, and
    logmean's `-da / (logL - logR)` amplifies one ulp of a logarithm by 1 / |da / a| (up to 1e4), the near-cancelling sum of
    S_ij f_ij by another ~1e2: with np.log this restatement and the C++ oracle agree to 2e-11 on a developed Kelvin-Helmholtz
    state, with the same libm logarithm to 2e-13 (the reference formulation's own sensitivity)."""
    x = np.asarray(x, dtype=float)
    good = x > 0
    return np.where(good, _mlog(np.where(good, x, 1.0)).astype(float), np.where(x == 0, -np.inf, np.nan))   # (C's log: -inf at 0, NaN below)


def EXP(x):
    return _mexp(np.asarray(x, dtype=float)).astype(float)


def POW(x, y):
    """libm's pow; a negative base gives NaN (math.pow would raise), as in the oracle and the CUDA path."""
    x = np.asarray(x, dtype=float)
    with np.errstate(invalid="ignore"):
        safe = np.where(x < 0, np.nan, x)
    return _mpow(safe, y).astype(float)


def flux_options(rhs):
    """(low-order surface flux, high-order surface flux, volume flux) codes of a LowOrderPositivity / FluxDiffRHS / LimitedDG."""
    if rhs.code == T.RHS_LOW_ORDER_POSITIVITY:
        return rhs.surface_flux.code, T.SURFFLUX_LF_PROJECTED, T.VOLFLUX_CHANDRASHEKAR
    if rhs.code == T.RHS_FLUX_DIFF:
        return T.SURFFLUX_LF_NODAL, rhs.surface_flux.code, rhs.volume_flux.code
    return rhs.low_order_surface_flux.code, rhs.high_order_surface_flux.code, rhs.high_order_volume_flux.code


# ---- compressible_Navier_Stokes.jl (Dim2); U[..., 4]
def pfun(g, U):
    return (g - 1.0) * (U[..., 3] - 0.5 * (U[..., 1] ** 2 + U[..., 2] ** 2) / U[..., 0])


def betafun(g, U):
    return U[..., 0] / (2 * pfun(g, U))


def wavespeed(g, U, n):
    """wavespeed_estimate(::Dim2, U, n) :48-62: the 1D estimate of (rho, n . m, E)."""
    mn = n[..., 0] * U[..., 1] + n[..., 1] * U[..., 2]
    p = (g - 1.0) * (U[..., 3] - 0.5 * mn ** 2 / U[..., 0])
    return np.abs(mn / U[..., 0]) + np.sqrt(g * p / U[..., 0])


def v_ufun(g, U):
    p = pfun(g, U)
    s = LOG(p / POW(U[..., 0], g))
    return np.stack([(g + 1 - s) - (g - 1) * U[..., 3] / p, U[..., 1] * (g - 1) / p, U[..., 2] * (g - 1) / p,
                     -U[..., 0] * (g - 1) / p], axis=-1)


def u_vfun(g, V):
    q = V[..., 1] ** 2 + V[..., 2] ** 2
    s = g - V[..., 0] + q / (2 * V[..., 3])
    rhoeV = POW((g - 1) / POW(-V[..., 3], g), 1 / (g - 1)) * EXP(-s / (g - 1))
    return np.stack([-rhoeV * V[..., 3], rhoeV * V[..., 1], rhoeV * V[..., 2], rhoeV * (1 - q / (2 * V[..., 3]))], axis=-1)


def fluxes(g, U):
    """-> [..., 2, 4]"""
    rho, m1, m2, E = (U[..., c] for c in range(4))
    p = pfun(g, U)
    u, v = m1 / rho, m2 / rho
    fx = np.stack([m1, m1 * u + p, rho * u * v, u * (E + p)], axis=-1)
    fy = np.stack([m2, rho * u * v, m2 * v + p, v * (E + p)], axis=-1)
    return np.stack([fx, fy], axis=-2)


def logmean(aL, aR, logL, logR):
    da = aR - aL
    aavg = 0.5 * (aR + aL)
    f = da / aavg
    v = f * f
    series = aavg * (1 + v * (-0.2 - v * (0.0512 - v * 0.026038857142857)))
    with np.errstate(divide="ignore", invalid="ignore"):
        general = -da / (logL - logR)
    return np.where(np.abs(f) < 1e-4, series, general)


def fS(g, rhoL, uL, vL, bL, rlL, blL, rhoR, uR, vR, bR, rlR, blR):
    """fS(::Dim2) :238-266 -> [..., 2, 4]"""
    rholog = logmean(rhoL, rhoR, rlL, rlR)
    betalog = logmean(bL, bR, blL, blR)
    rhoavg, uavg, vavg = 0.5 * (rhoL + rhoR), 0.5 * (uL + uR), 0.5 * (vL + vR)
    unorm = uL * uR + vL * vR
    pa = rhoavg / (bL + bR)
    f4aux = rholog / (2 * (g - 1) * betalog) + pa + 0.5 * rholog * unorm
    Fx1 = rholog * uavg
    Fx3 = Fx1 * vavg
    Fy1 = rholog * vavg
    fx = np.stack([Fx1, Fx1 * uavg + pa, Fx3, f4aux * uavg], axis=-1)
    fy = np.stack([Fy1, Fx3, Fy1 * vavg + pa, f4aux * vavg], axis=-1)
    return np.stack([fx, fy], axis=-2)


def find_alpha(POSTOL, ui, ut):
    """low_order_graph_viscosity.jl:299-327, elementwise over leading axes."""
    def ok(al):
        s = al[..., None] * ui - ut
        with np.errstate(divide="ignore", invalid="ignore"):
            rhoe = s[..., 3] - 0.5 * (s[..., 1] ** 2 + s[..., 2] ** 2) / s[..., 0]
        return (s[..., 0] > POSTOL) & (rhoe > POSTOL)
    aL = np.zeros(ui.shape[:-1])
    aR = np.ones(ui.shape[:-1])
    for _ in range(1100):
        bad = ~ok(aR)
        if not bad.any():
            break
        aR = np.where(bad, 2 * aR, aR)
    for _ in range(50):
        aM = (aL + aR) / 2
        good = ok(aM)
        aR = np.where(good, aM, aR)
        aL = np.where(good, aL, aM)
    return aR


def dense_theta(param, dd, Uq):
    """compute_entropyproj_limiting_param!(::GaussCollocation) with NodewiseScaledExtrapolation (filter.jl:6-130): per face node the
    largest theta in [0, 1] (21-step bisection, nonlinear_solvers.jl:3-20) for which u(v_tilde_f(theta)) stays within the bounds of
    :84-98; all face nodes of all elements at once.  -> theta_local [K, Nfp]"""
    g = param.equation.gamma
    ops = dd.ops
    eps, zeta, eta = param.global_constants.POSTOL, param.limiting_param.zeta, param.limiting_param.eta
    vq = v_ufun(g, Uq)
    Uf = np.einsum("fq,kqc->kfc", ops.Vf, Uq)                                  # calc_face_values! :26-41
    VUf = np.einsum("fq,kqc->kfc", ops.Vf, vq)
    rhoef = rhoe_ufun(Uf)
    hi, lo = np.einsum("fq,kqc->kfc", ops.Vf, vq), np.einsum("fq,kqc->kfc", ops.Vf_low, vq)

    def ok(th):
        # (theta Vf + (1 - theta) Vf_low) vq accumulated node by node in the reference; the two extrapolations are linear in vq
        W = th[..., None] * np.asarray(ops.Vf)[None] + (1 - th[..., None]) * np.asarray(ops.Vf_low)[None]      # [K, Nfp, Nq]
        vt = np.einsum("kfq,kqc->kfc", W, vq)
        well = vt[..., 3] < -eps
        with np.errstate(all="ignore"):
            ut = u_vfun(g, np.where(well[..., None], vt, np.array([0.0, 0.0, 0.0, -1.0])))
            rhoe = rhoe_ufun(ut)
            good = (vt[..., 3] < np.minimum(zeta * VUf[..., 3], -eps)) & (ut[..., 0] > np.maximum((1 - eta) * Uf[..., 0], eps)) & \
                (ut[..., 0] < (1 + eta) * Uf[..., 0]) & (rhoe > np.maximum((1 - eta) * rhoef, eps)) & (rhoe < (1 + eta) * rhoef)
        return well & good
    one = ok(np.ones(Uf.shape[:2]))
    xv, xi = np.zeros(Uf.shape[:2]), np.ones(Uf.shape[:2])
    for _ in range(21):
        xn = 0.5 * (xv + xi)
        good = ok(xn)
        xv = np.where(good, xn, xv)
        xi = np.where(good, xi, xn)
    return np.where(one, 1.0, xv)


def dense_rhs(param, dd, bc, Uq, t, nstage=1, theta_local=None):
    """One rhs! without the limiter.  Returns dict(rhsL, rhsH, rhsxyL, rhsxyH [K, Nq, 2, 4], dt, u_tilde_f, ...).
    `theta_local` [K, Nfp]: the projection-limiting parameters to use (NodewiseScaledExtrapolation); None = computed here when
    the configuration has that limiter, 1 otherwise."""
    g = param.equation.gamma
    sz, ops, geom = dd.sizes, dd.ops, dd.geom
    K, Nq, Nfp, Nh = sz.K, sz.Nq, sz.Nfp, sz.Nh
    N1D = param.N + 1
    POSTOL = param.global_constants.POSTOL
    tp = param.timestepping_param
    fq2q = np.asarray(ops.fq2q) - 1
    rxJ, sxJ, ryJ, syJ = (np.asarray(a, dtype=float) for a in geom.GJh)        # [K, Nh]
    low_flux, high_flux, vol_flux = flux_options(param.rhs)
    low_proj = low_flux == T.SURFFLUX_LF_PROJECTED
    mapP = np.asarray(bc.mapP).reshape(K * Nfp) - 1                            # linear index into [K, Nfp]
    mapI = np.asarray(bc.mapI, dtype=np.int64).reshape(-1) - 1
    mapO = np.asarray(bc.mapO, dtype=np.int64).reshape(-1) - 1
    Ival = np.asarray(bc.Ival, dtype=float).reshape(-1, 4)

    # ---- entropy projection, theta = 1 (rhs.jl:59-133)
    vq = v_ufun(g, Uq)                                                         # [K, Nq, 4]
    if theta_local is None:
        theta_local = dense_theta(param, dd, Uq) if param.entropyproj_limiter.code == T.PROJLIM_NODEWISE else np.ones((K, Nfp))
    Vf_new = theta_local[..., None] * np.asarray(ops.Vf)[None] + (1 - theta_local[..., None]) * np.asarray(ops.Vf_low)[None]   # [K, Nfp, Nq]
    vf = np.einsum("kfq,kqc->kfc", Vf_new, vq)                                 # entropy_projection_face_node! :84-94
    utf = u_vfun(g, vf)                                                        # [K, Nfp, 4]
    u_tilde = np.concatenate([Uq, utf], axis=1)                                # [K, Nh, 4]

    # physical boundary matrix entries and normals at the face nodes (rhs_utils.jl:13-39)
    Br, Bs = (np.asarray(b, dtype=float) for b in ops.Brs)
    Bxy = np.stack([rxJ[:, Nq:] * Br + sxJ[:, Nq:] * Bs, ryJ[:, Nq:] * Br + syJ[:, Nq:] * Bs], axis=-1)   # [K, Nfp, 2]
    nnorm = np.linalg.norm(Bxy, axis=-1)
    nf = Bxy / nnorm[..., None]
    xface = np.arange(Nfp) < 2 * N1D                                           # apply_LF_dissipation_to_BF :84-91

    # =================== low order (low_order_graph_viscosity.jl)
    Uf = utf if low_proj else Uq[:, fq2q]                                      # update_face_values! :44-65
    flux_q = fluxes(g, Uq)                                                     # [K, Nq, 2, 4]
    flux_f = fluxes(g, Uf)
    ws_f = wavespeed(g, Uf, nf)
    uP = Uf.reshape(K * Nfp, 4)[mapP].copy()                                   # get_uP_and_enforce_BC! :87-118
    if len(mapI):
        uP[mapI] = Ival
    if len(mapO):
        uP[mapO] = Uq[mapO // Nfp, fq2q[mapO % Nfp]]
    uP = uP.reshape(K, Nfp, 4)
    Q0F1 = np.zeros((K, Nq, 2, 4))
    lam = np.zeros((K, Nq, Nq))
    Sr0, Ss0 = ops.Srs0
    for (i1, j1) in ops.Srs0_nnz:                                              # accumulate_low_order_rhs_volume! :131-165
        i, j = i1 - 1, j1 - 1
        Sx = rxJ[:, i] * Sr0[i, j] + sxJ[:, i] * Ss0[i, j]
        Sy = ryJ[:, i] * Sr0[i, j] + syJ[:, i] * Ss0[i, j]
        Sxy = np.stack([Sx, Sy], axis=-1)                                      # [K, 2]
        nn = np.linalg.norm(Sxy, axis=-1)
        nij = Sxy / nn[:, None]
        lam_ij = nn * np.maximum(wavespeed(g, Uq[:, i], nij), wavespeed(g, Uq[:, j], -nij))
        lam[:, i, j] = lam[:, j, i] = lam_ij
        F = 0.5 * (flux_q[:, i] + flux_q[:, j])                                # [K, 2, 4]
        visc = lam_ij[:, None] * (Uq[:, j] - Uq[:, i])                         # graph_viscosity(::Dim2) :118-132
        D = np.zeros((K, 2, 4))
        isx = np.abs(Sx) > 1e-10
        D[isx, 0] = visc[isx]
        D[~isx, 1] = visc[~isx]
        SF = 2.0 * Sxy[:, :, None] * F - D
        Q0F1[:, i] += SF
        Q0F1[:, j] -= SF
    rhsxyL = -Q0F1
    lamB = 0.5 * nnorm * np.maximum(ws_f, ws_f.reshape(K * Nfp)[mapP].reshape(K, Nfp))     # :167-195
    fstar_L = 0.5 * (flux_f + fluxes(g, uP))
    BF_L = Bxy[..., None] * fstar_L                                            # [K, Nfp, 2, 4]
    lf = lamB[..., None] * (uP - Uf)
    fstar_L = fstar_L.copy()                                                   # apply_LF_dissipation_to_fstar, rhs_utils.jl:93-102
    fstar_L[:, xface, 0] -= lf[:, xface] / Bxy[:, xface, 0:1]
    fstar_L[:, ~xface, 1] -= lf[:, ~xface] / Bxy[:, ~xface, 1:2]
    BF_L[:, xface, 0] -= lf[:, xface]
    BF_L[:, ~xface, 1] -= lf[:, ~xface]
    for f in range(Nfp):
        rhsxyL[:, fq2q[f]] -= BF_L[:, f]
    wJ = np.asarray(geom.Jq, dtype=float) * ops.wq[None, :]                    # scale_low_order_rhs_by_mass! :197-211
    rhsxyL = rhsxyL / wJ[:, :, None, None]
    rhsL = rhsxyL.sum(axis=2)
    dt = None
    if nstage == 1:                                                            # calculate_lambda_and_low_order_CFL! :213-281
        lam_i = lam.sum(axis=2)
        if low_proj:
            alpha = find_alpha(POSTOL, Uq[:, fq2q], utf)
            lamB_cfl = alpha * lamB + 0.5 * nnorm * ws_f
        else:
            lamB_cfl = lamB
        for i in range(Nq):
            for f1 in ops.q2fq[i]:
                lam_i[:, i] += lamB_cfl[:, f1 - 1]
        dt = min(min(tp.CFL * tp.dt0, tp.T - t), float((tp.CFL * 0.5 * wJ / lam_i).min()))

    # =================== high order (flux_differencing.jl)
    beta = betafun(g, u_tilde)                                                 # calculate_primitive_variables! :39-72
    rholog, betalog = LOG(u_tilde[..., 0]), LOG(beta)
    uu, vv = u_tilde[..., 1] / u_tilde[..., 0], u_tilde[..., 2] / u_tilde[..., 0]
    lam_f = wavespeed(g, utf, nf)                                              # calculate_interface_dissipation_coeff! :92-116
    LFc = 0.5 * nnorm * np.maximum(lam_f, lam_f.reshape(K * Nfp)[mapP].reshape(K, Nfp))
    uPh = utf.reshape(K * Nfp, 4)[mapP].copy()
    LFc = LFc.reshape(K * Nfp)
    if len(mapI):                                                              # enforce_BC! :118-151
        LFc[mapI] = 0.0
        uPh[mapI] = Ival
    if len(mapO):
        LFc[mapO] = 0.0
        uPh[mapO] = Uq[mapO // Nfp, fq2q[mapO % Nfp]]
    LFc = LFc.reshape(K, Nfp)
    uPh = uPh.reshape(K, Nfp, 4)
    Srh, Ssh = ops.Srsh_db
    Sxh = rxJ[:, :, None] * Srh[None] + sxJ[:, :, None] * Ssh[None]            # Sx(::Dim2) :49-54: geometry at row i
    Syh = ryJ[:, :, None] * Srh[None] + syJ[:, :, None] * Ssh[None]            # [K, Nh, Nh]
    vol_central = vol_flux == T.VOLFLUX_CENTRAL
    if vol_central:                                                            # eval_high_order_volume_flux(::CentralFlux) :217-221
        fh = fluxes(g, u_tilde)                                                # [K, Nh, 2, 4]
        Fij = 0.5 * (fh[:, :, None] + fh[:, None, :])                          # [K, Nh, Nh, 2, 4]
    else:
        L = lambda a: a[:, :, None]
        R = lambda a: a[:, None, :]
        Fij = fS(g, L(u_tilde[..., 0]), L(uu), L(vv), L(beta), L(rholog), L(betalog),
                 R(u_tilde[..., 0]), R(uu), R(vv), R(beta), R(rholog), R(betalog))
    QF1 = np.stack([np.einsum("kij,kijc->kic", Sxh, Fij[..., 0, :]), np.einsum("kij,kijc->kic", Syh, Fij[..., 1, :])], axis=2)   # [K, Nh, 2, 4]
    surf_chandra = high_flux == T.SURFFLUX_CHANDRASHEKAR_PROJECTED
    if surf_chandra:                                                           # ChandrashekarOnProjectedVal :232-243
        bP = betafun(g, uPh)
        fstar_H = fS(g, utf[..., 0], utf[..., 1] / utf[..., 0], utf[..., 2] / utf[..., 0], beta[:, Nq:], rholog[:, Nq:], betalog[:, Nq:],
                     uPh[..., 0], uPh[..., 1] / uPh[..., 0], uPh[..., 2] / uPh[..., 0], bP, LOG(uPh[..., 0]), LOG(bP))
    else:
        fstar_H = 0.5 * (fluxes(g, utf) + fluxes(g, uPh))
    BF_H = Bxy[..., None] * fstar_H
    lfh = LFc[..., None] * (uPh - utf)
    fstar_H = fstar_H.copy()
    fstar_H[:, xface, 0] -= lfh[:, xface] / Bxy[:, xface, 0:1]
    fstar_H[:, ~xface, 1] -= lfh[:, ~xface] / Bxy[:, ~xface, 1:2]
    BF_H[:, xface, 0] -= lfh[:, xface]
    BF_H[:, ~xface, 1] -= lfh[:, ~xface]
    if (theta_local == 1.0).all():
        proj = np.einsum("qh,khdc->kqdc", ops.MinvVhT, QF1) + np.einsum("qf,kfdc->kqdc", ops.MinvVfT, BF_H)
    else:   # project_flux_difference_to_quad!(::ScaledExtrapolation) :288-319: M^-1 [I Vf_new^T] with the limited Vf_new
        proj = (QF1[:, :Nq] + np.einsum("kfq,kfdc->kqdc", Vf_new, QF1[:, Nq:]) + np.einsum("kfq,kfdc->kqdc", Vf_new, BF_H)) \
            / np.asarray(ops.wq)[None, :, None, None]
    rhsxyH = -proj / np.asarray(geom.Jq, dtype=float)[:, :, None, None]        # assemble_rhs! :331-361
    return {"rhsL": rhsL, "rhsxyL": rhsxyL, "rhsH": rhsxyH.sum(axis=2), "rhsxyH": rhsxyH, "dt": dt, "u_tilde_f": utf,
            "BF_L": BF_L, "BF_H": BF_H, "wJ": wJ, "theta_local": theta_local, "fstar_L": fstar_L, "fstar_H": fstar_H}


# ---- limiter_utils.jl:26-95, vectorised (IEEE semantics kept: a == 0 gives infinite / NaN roots, which fail every comparison)
def rhoe_quadratic_solve(ZEROTOL, U, Pv, Lrhoe):
    a = Pv[..., 0] * Pv[..., 3] - 0.5 * (Pv[..., 1] ** 2 + Pv[..., 2] ** 2)
    b = U[..., 3] * Pv[..., 0] + U[..., 0] * Pv[..., 3] - U[..., 1] * Pv[..., 1] - U[..., 2] * Pv[..., 2] - Pv[..., 0] * Lrhoe
    c = U[..., 3] * U[..., 0] - 0.5 * (U[..., 1] ** 2 + U[..., 2] ** 2) - U[..., 0] * Lrhoe
    disc = b * b - 4 * a * c
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.sqrt(np.where(disc >= 0, disc, 0.0))
        r1 = (-b + sq) / (2 * a)
        r2 = (-b - sq) / (2 * a)
        l = np.ones_like(a)
        both = (r1 > ZEROTOL) & (r2 > ZEROTOL)
        only1 = ~both & (r1 > ZEROTOL) & (r2 < -ZEROTOL)
        only2 = ~both & ~only1 & (r2 > ZEROTOL) & (r1 < -ZEROTOL)
        l = np.where(both, np.minimum(r1, r2), l)       # (finite here: both compare greater than ZEROTOL)
        l = np.where(only1, r1, l)
        l = np.where(only2, r2, l)
    return np.where(disc >= 0, l, 1.0)


def rhoe_ufun(U):
    return U[..., 3] - 0.5 * (U[..., 1] ** 2 + U[..., 2] ** 2) / U[..., 0]


def limiting_param_bounds(ZEROTOL, U, Pv, Lrho, Lrhoe, Urho=None):
    """limiting_param_bound_rho_rhoe (limiter_utils.jl:26-40) with an optional finite upper density bound (TVD bounds); Urhoe = Inf."""
    with np.errstate(divide="ignore", invalid="ignore"):
        l = np.where(U[..., 0] + Pv[..., 0] < Lrho, np.maximum((Lrho - U[..., 0]) / Pv[..., 0], 0.0), 1.0)
        if Urho is not None:
            l = np.where(U[..., 0] + Pv[..., 0] > Urho, np.minimum(l, np.maximum((Urho - U[..., 0]) / Pv[..., 0], 0.0)), l)
    return np.minimum(np.minimum(l, rhoe_quadratic_solve(ZEROTOL, U, Pv, Lrhoe)), 1.0)


def s_modified(g, U):
    """s_modified_ufun :80-85 (a negative density gives NaN, which fails the bound test)."""
    return rhoe_ufun(U) * POW(U[..., 0], -g)


def limiting_param_phi(g, POSTOL, U, Pv, Lphi, lpos):
    """limiting_param_bound_phi (limiter_utils.jl:42-50): bisection(f, 0, lpos), nonlinear_solvers.jl:3-20, all entries at once."""
    def f(l):
        with np.errstate(all="ignore"):
            return s_modified(g, U + l[..., None] * Pv) >= Lphi - POSTOL
    top = f(lpos)
    xv, xi = np.zeros_like(lpos), lpos.copy()
    for _ in range(21):
        xn = 0.5 * (xv + xi)
        good = f(xn)
        xv = np.where(good, xn, xv)
        xi = np.where(good, xi, xn)
    return np.where(top, lpos, xv)


def smoothness_indicator(param, dd, Uq):
    """initialize_smoothness_indicator!(::Dim2) shock_capture.jl:47-94: modal energies of rho p in the two highest modes."""
    g, N = param.equation.gamma, param.N
    modal = np.einsum("mq,kq->km", dd.ops.VDM_inv, Uq[..., 0] * pfun(g, Uq))
    e = (modal ** 2).reshape(-1, N + 1, N + 1)                                 # [K, j, i]
    ii, jj = np.meshgrid(np.arange(N + 1), np.arange(N + 1))
    eN = np.where((ii == N) | (jj == N), e, 0.0).reshape(len(e), -1)
    eNm1 = np.where((ii == N - 1) | (jj == N - 1), e, 0.0).reshape(len(e), -1)
    tot = e.reshape(len(e), -1)
    sN, sNm1, st = np.zeros(len(e)), np.zeros(len(e)), np.zeros(len(e))
    for m in range(tot.shape[1]):                                              # the reference's accumulation order
        sN, sNm1, st = sN + eN[:, m], sNm1 + eNm1[:, m], st + tot[:, m]
    return np.maximum(sN / st, sNm1 / st)


def stencil_neighbours(dd, bc, n):
    """low_order_stencil(::Dim2) limiter_utils.jl:222-231 for every node: (k, node) of the left / right / bottom / top stencil
    node, across element faces through q2fq, mapP and fq2q (quad_index_to_quad_index_P :197-210).  -> kN, qN [K, Nq, 4]"""
    K, Nq, Nfp = dd.sizes.K, dd.sizes.Nq, dd.sizes.Nfp
    mapP = np.asarray(bc.mapP).reshape(K, Nfp) - 1
    fq2q = np.asarray(dd.ops.fq2q) - 1
    kN = np.empty((K, Nq, 4), dtype=np.int64)
    qN = np.empty((K, Nq, 4), dtype=np.int64)
    kk = np.arange(K)
    for q in range(Nq):
        i, j = q % n, q // n
        faces = [f - 1 for f in dd.ops.q2fq[q]]
        for s_, (inside, qin, direction) in enumerate(((i - 1 >= 0, q - 1, 0), (i + 1 <= n - 1, q + 1, 0),
                                                       (j - 1 >= 0, q - n, 1), (j + 1 <= n - 1, q + n, 1))):
            if inside:
                kN[:, q, s_], qN[:, q, s_] = kk, qin
            else:
                f = faces[0] if len(faces) == 1 else faces[direction]
                P = mapP[:, f]
                kN[:, q, s_], qN[:, q, s_] = P // Nfp, fq2q[P % Nfp]
    return kN, qN


def smooth_factor(param, sigma):
    """update_smoothness_factor! of the relaxed bounds, subcell.jl:937-956."""
    s0, sk = math.log10(float(param.N) ** -4), np.log10(sigma)
    return np.where(sk < s0 - 1.0, 0.0, np.where(sk > s0 + 1.0, 1.0, 0.5 - 0.5 * np.sin(math.pi * (sk - s0) / 2.0)))


def es_volume(param, dd, Uq, d, fb, Lx, Ly, epsk, bound):
    """enforce_ES_subcell! on Lobatto nodes = initialize_ES_subcell_limiting! + enforce_ES_subcell_volume! (subcell.jl:508-565,
    630-743; the interface part is a no-op there, :755-757): per element and direction, if the entropy production estimate of the
    positivity-limited interior subcell fluxes exceeds its budget, the faces with the largest production are switched off in
    descending (value, index) order, the last one partially.  Element by element (the greedy step is inherently serial); Lx, Ly are
    updated in place."""
    g, ZEROTOL = param.equation.gamma, param.global_constants.ZEROTOL
    K, Nq, Nfp = dd.sizes.K, dd.sizes.Nq, dd.sizes.Nfp
    n = param.N + 1
    ops, geom = dd.ops, dd.geom
    fq2q = np.asarray(ops.fq2q) - 1
    rxJ, sxJ, ryJ, syJ = (np.asarray(a, dtype=float) for a in geom.GJh)
    Br, Bs = (np.asarray(b, dtype=float) for b in ops.Brs)
    Bxy = np.stack([rxJ[:, Nq:] * Br + sxJ[:, Nq:] * Bs, ryJ[:, Nq:] * Br + syJ[:, Nq:] * Bs], axis=-1)
    vq = v_ufun(g, Uq).reshape(K, n, n, 4)                                     # [K, j, i]
    psif = (g - 1.0) * Uq[:, fq2q][..., 1:3]                                   # psi_ufun :126-129 at the face nodes' volume nodes
    relaxed = bound.code in (T.BOUND_POS_RELAXED_CELL_ENTROPY, T.BOUND_TVD_RELAXED_CELL_ENTROPY)
    for k in range(K):
        sB = np.zeros(2)
        for f in range(Nfp):
            sB = sB + Bxy[k, f] * psif[k, f]
        for dirn in (0, 1):
            fH, fL = fb["H"][dirn][k], fb["L"][dirn][k]                        # x: [sj, si, 4]; y: [sj, si, 4]
            L = Lx[k] if dirn == 0 else Ly[k]
            faces = []                                                         # (flat index in the reference's dvdf order, sj, si)
            if dirn == 0:
                for sj in range(n):
                    for si in range(1, n):
                        faces.append(((si - 1) + sj * (n - 1), sj, si, vq[k, sj, si - 1] - vq[k, sj, si]))
                visit = faces                                                  # sums run sj outer, si inner (:653-658)
            else:
                for sj in range(1, n):
                    for si in range(n):
                        faces.append((si + (sj - 1) * n, sj, si, vq[k, sj - 1, si] - vq[k, sj, si]))
                visit = sorted(faces, key=lambda r: (r[2], r[1]))              # sums run si outer, sj inner (:663-668)
            dvdf = {r[0]: float(np.sum(r[3] * (fH[r[1], r[2]] - fL[r[1], r[2]]))) for r in faces}
            sdvfL, spos = 0.0, 0.0
            for r in visit:
                sdvfL += float(np.sum(r[3] * fL[r[1], r[2]]))
            for r in visit:
                spos += L[r[1], r[2]] * dvdf[r[0]]
            budget = sB[dirn] - sdvfL
            rhs_ = (1 - bound.beta * epsk[k]) * budget if relaxed else budget  # rhs_es :746-753
            tol = max(0.0, sdvfL - sB[dirn])
            if not (spos - rhs_ > tol):
                continue
            where = {r[0]: (r[1], r[2]) for r in faces}
            order = sorted(dvdf, key=lambda e: (dvdf[e], e), reverse=True)     # sort!(..., rev=true) on (value, index) tuples
            lhs, taken = spos, []
            for e in order:
                if not lhs > rhs_ + tol:
                    break
                if dvdf[e] < ZEROTOL:
                    break
                lhs = lhs - L[where[e]] * dvdf[e]
                taken.append(e)
            for m, e in enumerate(taken):
                l_new = max((rhs_ + tol - lhs) / dvdf[e], 0.0) if m == len(taken) - 1 else 0.0
                L[where[e]] = min(L[where[e]], l_new)


def es_interface(param, dd, bc, Uq, d, Lx, Ly):
    """enforce_ES_subcell_interface!(::Dim2, ::GaussCollocation) subcell.jl:759-823, in element order (what one thread of the
    reference does; with several threads its result depends on their interleaving, oracle deviation D5): per element face node,
    the largest l <= min(own, partner's CURRENT coefficient) with l dv.f*_H + (1 - l) dv.f*_L <= dpsi by the 21-step bisection; only
    the own entry is written.  Lx, Ly updated in place."""
    g = param.equation.gamma
    K, Nq, Nfp = dd.sizes.K, dd.sizes.Nq, dd.sizes.Nfp
    n = param.N + 1
    fq2q = np.asarray(dd.ops.fq2q) - 1
    mapP = np.asarray(bc.mapP).reshape(K, Nfp) - 1
    vf = v_ufun(g, Uq[:, fq2q])                                                # at the face nodes' volume nodes (:508-530)
    psif = (g - 1.0) * Uq[:, fq2q][..., 1:3]
    fH, fL = d["fstar_H"], d["fstar_L"]                                        # [K, Nfp, 2, 4]

    def solve(l, dvfH, dvfL, dpsi):                                            # solve_l_es_interface! :818-823
        ok = lambda x: x * dvfH + (1 - x) * dvfL <= dpsi
        if ok(l):
            return l
        xv, xi = 0.0, l
        for _ in range(21):
            xn = 0.5 * (xv + xi)
            if ok(xn):
                xv = xn
            else:
                xi = xn
        return xv
    for k in range(K):
        for dirn, L in ((0, Lx), (1, Ly)):
            for line in range(n):
                for e, pos in ((0, 0), (1, n)):
                    f = (2 * dirn + e) * n + line
                    P = mapP[k, f]
                    kP, iP = P // Nfp, P % Nfp
                    if dirn == 0:
                        own, partner = (k, line, pos), (kP, iP % n, 0 if iP // n == 0 else n)
                    else:
                        own, partner = (k, pos, line), (kP, 0 if iP // n == 2 else n, iP % n)
                    dv = vf[k, f] - vf[kP, iP]
                    dpsi = psif[k, f, dirn] - psif[kP, iP, dirn]
                    dvfH, dvfL = float(np.sum(dv * fH[k, f, dirn])), float(np.sum(dv * fL[k, f, dirn]))
                    L[own] = solve(min(L[own], L[partner]), dvfH, dvfL, dpsi)


def dense_limited_rhs(param, dd, bc, Uq, t, dt, nstage=1, theta_local=None, smin=None):
    """rhs!(::LimitedDG): dense_rhs + apply_rhs_limiter! (limiter.jl:8-56) -- Zhang-Shu (zhangshu.jl:4-45) or the subcell limiter
    (subcell.jl:4-349, 418-456, 841-924) with PositivityBound, the minimum-entropy bounds (plain / relaxed), the TVD bounds and their
    combinations, with or without HennemannShockCapture (shock_capture.jl) -- as whole-array operations, neighbours through mapP.
    `dt` is the dt the limiter sees (the caller's, rhs.jl:46,52); `smin` the minimum of s_modified over the initial condition
    (subcell.jl:31-34: recorded at t == t0, nstage == 1).  Cell-entropy bounds: Lobatto nodes only (`es_volume`).
    Adds rhsU and L [K] or Lx [K, N1D, N1D+1], Ly [K, N1D+1, N1D]."""
    d = dense_rhs(param, dd, bc, Uq, t, nstage, theta_local)
    g = param.equation.gamma
    zeta, ZEROTOL, POSTOL = param.limiting_param.zeta, param.global_constants.ZEROTOL, param.global_constants.POSTOL
    sz = dd.sizes
    K, Nq, Nfp = sz.K, sz.Nq, sz.Nfp
    n = param.N + 1
    lim = param.rhs_limiter
    subcell = lim.code == T.LIMITER_SUBCELL
    bcode = lim.bound.code if subcell else T.BOUND_POSITIVITY
    cell = bcode in (T.BOUND_POS_CELL_ENTROPY, T.BOUND_POS_RELAXED_CELL_ENTROPY, T.BOUND_TVD_CELL_ENTROPY, T.BOUND_TVD_RELAXED_CELL_ENTROPY)
    relaxed_cell = bcode in (T.BOUND_POS_RELAXED_CELL_ENTROPY, T.BOUND_TVD_RELAXED_CELL_ENTROPY)
    tvd = bcode >= T.BOUND_TVD
    minent = bcode in (T.BOUND_POS_MIN_ENTROPY, T.BOUND_POS_RELAXED_MIN_ENTROPY, T.BOUND_TVD_MIN_ENTROPY, T.BOUND_TVD_RELAXED_MIN_ENTROPY)
    relaxed = bcode in (T.BOUND_POS_RELAXED_MIN_ENTROPY, T.BOUND_TVD_RELAXED_MIN_ENTROPY)
    hen = lim.shockcapture.code == T.SHOCKCAPTURE_HENNEMANN
    blend = np.ones(K)
    if hen or relaxed or relaxed_cell:
        sigma = smoothness_indicator(param, dd, Uq)
    if hen:                                                                    # update_blending_factor! shock_capture.jl:111-132
        TN = lim.shockcapture.a * 10 ** (-lim.shockcapture.c * (param.N + 1) ** 0.25)
        s_factor = math.log((1 - 0.0001) / 0.0001)
        blend = np.maximum(np.minimum(1.0 - 1.0 / (1.0 + EXP(-s_factor / TN * (sigma - TN))), 1.0), 0.5)
    d["blending_factor"] = blend
    uL = Uq + dt * d["rhsL"]
    if not subcell:
        Pv = dt * (d["rhsH"] - d["rhsL"])
        L = limiting_param_bounds(ZEROTOL, uL, Pv, zeta * uL[..., 0], zeta * rhoe_ufun(uL)).min(axis=1)
        l = np.minimum(L, blend)
        d["L"] = L
        d["rhsU"] = (1 - l)[:, None, None] * d["rhsL"] + l[:, None, None] * d["rhsH"]
        return d
    if minent or tvd:
        kN, qN = stencil_neighbours(dd, bc, n)
    Lphi = None
    if minent:                                                                 # initialize_entropy_bounds! subcell.jl:14-75
        sm = s_modified(g, Uq)
        lb = np.minimum(sm, sm[kN, qN].min(axis=2))
        epsk = smooth_factor(param, sigma) if relaxed else np.ones(K)
        Lphi = (epsk[:, None] * lb + (1 - epsk[:, None]) * smin).reshape(K, n, n)
        d["lbound_s_modified"] = Lphi.reshape(K, Nq)
    wJ = d["wJ"].reshape(K, n, n)                                     # [K, jq, iq]
    uLg = uL.reshape(K, n, n, 4)
    Lrho, Urho = zeta * uLg[..., 0], None
    if tvd:                                                                    # initialize_TVD_bounds! :112-141
        rhoL = Uq[..., 0] + dt * d["rhsL"][..., 0]
        nb = rhoL[kN, qN]
        Lrho = np.minimum(rhoL, nb.min(axis=2)).reshape(K, n, n)
        Urho = np.maximum(rhoL, nb.max(axis=2)).reshape(K, n, n)
    Lrhoe = zeta * rhoe_ufun(uLg)

    def coef(Pv):                                                              # limiting_param limiter_utils.jl:4-24
        l = limiting_param_bounds(ZEROTOL, uLg, Pv, Lrho, Lrhoe, Urho)
        if minent:
            l = limiting_param_phi(g, POSTOL, uLg, Pv, Lphi, l)
        return l
    rH, rL = d["rhsxyH"].reshape(K, n, n, 2, 4), d["rhsxyL"].reshape(K, n, n, 2, 4)
    fb = {}
    for name, r, BF in (("H", rH, d["BF_H"]), ("L", rL, d["BF_L"])):  # accumulate_f_bar! :163-206 (running sums, in this order)
        fx = np.zeros((K, n, n + 1, 4))                               # [K, sj, si]
        fx[:, :, 0] = BF[:, 0:n, 0]
        for si in range(1, n + 1):
            fx[:, :, si] = fx[:, :, si - 1] + wJ[:, :, si - 1, None] * r[:, :, si - 1, 0]
        fy = np.zeros((K, n + 1, n, 4))                               # [K, sj, si]
        fy[:, 0] = BF[:, 2 * n:3 * n, 1]
        for sj in range(1, n + 1):
            fy[:, sj] = fy[:, sj - 1] + wJ[:, sj - 1, :, None] * r[:, sj - 1, :, 1]
        fb[name] = (fx, fy)
    dfx, dfy = fb["H"][0] - fb["L"][0], fb["H"][1] - fb["L"][1]
    Lx, Ly = np.ones((K, n, n + 1)), np.ones((K, n + 1, n))          # subcell_bound_limiter! :248-349
    # x: subcell face si is the LEFT face of node si (P = -4 dt df / wJ) and the RIGHT face of node si - 1 (P = +4 dt df / wJ)
    Lx[:, :, :n] = np.minimum(Lx[:, :, :n], coef(-4 * dt * dfx[:, :, :n] / wJ[..., None]))
    Lx[:, :, 1:] = np.minimum(Lx[:, :, 1:], coef(4 * dt * dfx[:, :, 1:] / wJ[..., None]))
    Ly[:, :n] = np.minimum(Ly[:, :n], coef(-4 * dt * dfy[:, :n] / wJ[..., None]))
    Ly[:, 1:] = np.minimum(Ly[:, 1:], coef(4 * dt * dfy[:, 1:] / wJ[..., None]))
    Lx, Ly = np.minimum(Lx, blend[:, None, None]), np.minimum(Ly, blend[:, None, None])       # "Apply shock capturing" :344-347
    if cell:
        es_volume(param, dd, Uq, d, fb, Lx, Ly, smooth_factor(param, sigma) if relaxed_cell else np.zeros(K), lim.bound)
        if param.approximation_basis.code == T.BASIS_GAUSS:
            es_interface(param, dd, bc, Uq, d, Lx, Ly)
    # symmetrize_limiting_parameters! :418-456, partner faces through mapP (limiter_utils.jl:122-181)
    mapP = np.asarray(bc.mapP).reshape(K, Nfp) - 1
    Lx0, Ly0 = Lx.copy(), Ly.copy()
    kk = np.arange(K)[:, None]
    line = np.arange(n)[None, :]
    for e, pos in ((0, 0), (1, n)):
        P = mapP[:, e * n:(e + 1) * n]                                # x faces: face nodes e n + sj
        kP, iP = P // Nfp, P % Nfp
        Lx[kk, line, pos] = np.minimum(Lx0[kk, line, pos], Lx0[kP, iP % n, np.where(iP // n == 0, 0, n)])
        P = mapP[:, (2 + e) * n:(3 + e) * n]                          # y faces: face nodes (2 + e) n + si
        kP, iP = P // Nfp, P % Nfp
        Ly[kk, pos, line] = np.minimum(Ly0[kk, pos, line], Ly0[kP, np.where(iP // n == 2, 0, n), iP % n])
    flx = Lx[..., None] * fb["H"][0] + (1 - Lx[..., None]) * fb["L"][0]     # accumulate_f_bar_limited! :841-876
    fly = Ly[..., None] * fb["H"][1] + (1 - Ly[..., None]) * fb["L"][1]
    rx = (flx[:, :, 1:] - flx[:, :, :-1]) / wJ[..., None]             # apply_subcell_limiter! :894-924
    ry = (fly[:, 1:] - fly[:, :-1]) / wJ[..., None]
    d["rhsU"] = (rx + ry).reshape(K, Nq, 4)
    d["Lx"], d["Ly"] = Lx, Ly
    return d


# ============================================================================================== 1D (Dim1), U[..., 3]
def p1(g, U):
    return (g - 1.0) * (U[..., 2] - 0.5 * U[..., 1] ** 2 / U[..., 0])


def ws1(g, U):
    return np.abs(U[..., 1] / U[..., 0]) + np.sqrt(g * p1(g, U) / U[..., 0])


def flux1(g, U):
    p = p1(g, U)
    u = U[..., 1] / U[..., 0]
    return np.stack([U[..., 1], U[..., 1] * u + p, u * (U[..., 2] + p)], axis=-1)


def v_u1(g, U):
    p = p1(g, U)
    s = LOG(p / POW(U[..., 0], g))
    return np.stack([(g + 1 - s) - (g - 1) * U[..., 2] / p, U[..., 1] * (g - 1) / p, -U[..., 0] * (g - 1) / p], axis=-1)


def u_v1(g, V):
    s = g - V[..., 0] + V[..., 1] ** 2 / (2 * V[..., 2])
    rhoeV = POW((g - 1) / POW(-V[..., 2], g), 1 / (g - 1)) * EXP(-s / (g - 1))
    return np.stack([-rhoeV * V[..., 2], rhoeV * V[..., 1], rhoeV * (1 - V[..., 1] ** 2 / (2 * V[..., 2]))], axis=-1)


def fS1(g, rhoL, uL, bL, rlL, blL, rhoR, uR, bR, rlR, blR):
    rholog, betalog = logmean(rhoL, rhoR, rlL, rlR), logmean(bL, bR, blL, blR)
    rhoavg, uavg = 0.5 * (rhoL + rhoR), 0.5 * (uL + uR)
    pa = rhoavg / (bL + bR)
    f4aux = rholog / (2 * (g - 1) * betalog) + pa + 0.5 * rholog * (uL * uR)
    F1 = rholog * uavg
    return np.stack([F1, F1 * uavg + pa, f4aux * uavg], axis=-1)


def rhoe1(U):
    return U[..., 2] - 0.5 * U[..., 1] ** 2 / U[..., 0]


def quad1(ZEROTOL, U, Pv, Lrhoe):
    """rhoe_quadratic_solve with rhoe_quadratic_coefficients(::Dim1), limiter_utils.jl:52-83."""
    a = Pv[..., 0] * Pv[..., 2] - 0.5 * Pv[..., 1] ** 2
    b = U[..., 2] * Pv[..., 0] + U[..., 0] * Pv[..., 2] - U[..., 1] * Pv[..., 1] - Pv[..., 0] * Lrhoe
    c = U[..., 2] * U[..., 0] - 0.5 * U[..., 1] ** 2 - U[..., 0] * Lrhoe
    disc = b * b - 4 * a * c
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.sqrt(np.where(disc >= 0, disc, 0.0))
        r1, r2 = (-b + sq) / (2 * a), (-b - sq) / (2 * a)
        both = (r1 > ZEROTOL) & (r2 > ZEROTOL)
        only1 = ~both & (r1 > ZEROTOL) & (r2 < -ZEROTOL)
        only2 = ~both & ~only1 & (r2 > ZEROTOL) & (r1 < -ZEROTOL)
        l = np.where(both, np.minimum(r1, r2), np.where(only1, r1, np.where(only2, r2, 1.0)))
    return np.where(disc >= 0, l, 1.0)


def lim1(ZEROTOL, U, Pv, Lrho, Lrhoe):
    with np.errstate(divide="ignore", invalid="ignore"):
        l = np.where(U[..., 0] + Pv[..., 0] < Lrho, np.maximum((Lrho - U[..., 0]) / Pv[..., 0], 0.0), 1.0)
    return np.minimum(np.minimum(l, quad1(ZEROTOL, U, Pv, Lrhoe)), 1.0)


def dense_theta_1d(param, dd, Uq):
    """NodewiseScaledExtrapolation on 1D Gauss nodes: filter.jl:6-130 is dimension-generic (Vf, Vf_low of the line element)."""
    g = param.equation.gamma
    ops = dd.ops
    eps, zeta, eta = param.global_constants.POSTOL, param.limiting_param.zeta, param.limiting_param.eta
    vq = v_u1(g, Uq)
    Uf = np.einsum("fq,kqc->kfc", ops.Vf, Uq)
    VUf = np.einsum("fq,kqc->kfc", ops.Vf, vq)
    rhoef = rhoe1(Uf)

    def ok(th):
        W = th[..., None] * np.asarray(ops.Vf)[None] + (1 - th[..., None]) * np.asarray(ops.Vf_low)[None]
        vt = np.einsum("kfq,kqc->kfc", W, vq)
        well = vt[..., 2] < -eps
        with np.errstate(all="ignore"):
            ut = u_v1(g, np.where(well[..., None], vt, np.array([0.0, 0.0, -1.0])))
            rhoe = rhoe1(ut)
            good = (vt[..., 2] < np.minimum(zeta * VUf[..., 2], -eps)) & (ut[..., 0] > np.maximum((1 - eta) * Uf[..., 0], eps)) & \
                (ut[..., 0] < (1 + eta) * Uf[..., 0]) & (rhoe > np.maximum((1 - eta) * rhoef, eps)) & (rhoe < (1 + eta) * rhoef)
        return well & good
    one = ok(np.ones(Uf.shape[:2]))
    xv, xi = np.zeros(Uf.shape[:2]), np.ones(Uf.shape[:2])
    for _ in range(21):
        xn = 0.5 * (xv + xi)
        good = ok(xn)
        xv, xi = np.where(good, xn, xv), np.where(good, xi, xn)
    return np.where(one, 1.0, xv)


def dense_limited_rhs_1d(param, dd, bc, Uq, t, dt, nstage=1, theta_local=None, smin=None):
    """rhs!(::LimitedDG) in 1D with NoEntropyProjectionLimiter, NoShockCapture, PositivityBound / Zhang-Shu: the same dense
    restatement for Dim1 (the Dim1 methods of the files cited above; Bx(::Dim1) as evidently intended, rhs_utils.jl:13-19 passes its
    arguments in the wrong order; the subcell symmetrisation pairs element k's first face with element k-1's last, periodically,
    whatever the boundary conditions, subcell.jl:405-416)."""
    g = param.equation.gamma
    sz, ops, geom = dd.sizes, dd.ops, dd.geom
    K, Nq, Nh = sz.K, sz.Nq, sz.Nh
    tp = param.timestepping_param
    zeta, ZEROTOL, POSTOL = param.limiting_param.zeta, param.global_constants.ZEROTOL, param.global_constants.POSTOL
    fq2q = np.asarray(ops.fq2q) - 1
    rxJ = np.asarray(geom.GJh[0], dtype=float)                                 # [K, Nh]
    low_flux, high_flux, vol_flux = flux_options(param.rhs)
    mapP = np.asarray(bc.mapP).reshape(K * 2) - 1
    mapI = np.asarray(bc.mapI, dtype=np.int64).reshape(-1) - 1
    mapO = np.asarray(bc.mapO, dtype=np.int64).reshape(-1) - 1
    Ival = np.asarray(bc.Ival, dtype=float).reshape(-1, 3)
    vq = v_u1(g, Uq)
    if theta_local is None:
        theta_local = dense_theta_1d(param, dd, Uq) if param.entropyproj_limiter.code == T.PROJLIM_NODEWISE else np.ones((K, 2))
    Vf_new = theta_local[..., None] * np.asarray(ops.Vf)[None] + (1 - theta_local[..., None]) * np.asarray(ops.Vf_low)[None]
    utf = u_v1(g, np.einsum("kfq,kqc->kfc", Vf_new, vq))
    u_tilde = np.concatenate([Uq, utf], axis=1)
    Bx = rxJ[:, Nq:] * np.asarray(ops.Brs[0], dtype=float)                      # [K, 2]
    nn = np.abs(Bx)
    # ---- low order
    Uf = utf if low_flux == T.SURFFLUX_LF_PROJECTED else Uq[:, fq2q]
    fq_, ff_ = flux1(g, Uq), flux1(g, Uf)
    wsf = ws1(g, Uf)
    uP = Uf.reshape(K * 2, 3)[mapP].copy()
    if len(mapI):
        uP[mapI] = Ival
    if len(mapO):
        uP[mapO] = Uq[mapO // 2, fq2q[mapO % 2]]
    uP = uP.reshape(K, 2, 3)
    S0 = ops.Srs0[0]
    Q0 = np.zeros((K, Nq, 3))
    lam = np.zeros((K, Nq, Nq))
    for (i1, j1) in ops.Srs0_nnz:
        i, j = i1 - 1, j1 - 1
        S = rxJ[:, i] * S0[i, j]
        lij = np.abs(S) * np.maximum(ws1(g, Uq[:, i]), ws1(g, Uq[:, j]))
        lam[:, i, j] = lam[:, j, i] = lij
        SF = 2.0 * S[:, None] * (0.5 * (fq_[:, i] + fq_[:, j])) - lij[:, None] * (Uq[:, j] - Uq[:, i])
        Q0[:, i] += SF
        Q0[:, j] -= SF
    rL = -Q0
    lamB = 0.5 * nn * np.maximum(wsf, wsf.reshape(K * 2)[mapP].reshape(K, 2))
    BF_L = Bx[..., None] * (0.5 * (ff_ + flux1(g, uP))) - lamB[..., None] * (uP - Uf)
    for f in range(2):
        rL[:, fq2q[f]] -= BF_L[:, f]
    wJ = np.asarray(geom.Jq, dtype=float) * np.asarray(ops.wq)[None, :]
    rhsL = rL / wJ[..., None]
    d = {"rhsL": rhsL}
    if nstage == 1:
        lam_i = lam.sum(axis=2)
        lamB_cfl = lamB
        if low_flux == T.SURFFLUX_LF_PROJECTED:
            ui, ut = Uq[:, fq2q], utf
            pad = lambda a: np.concatenate([a[..., :2], np.zeros_like(a[..., :1]), a[..., 2:]], axis=-1)   # (rho, m, 0, E) for the 2D helper
            lamB_cfl = find_alpha(POSTOL, pad(ui), pad(ut)) * lamB + 0.5 * nn * wsf
        for i in range(Nq):
            for f1 in ops.q2fq[i]:
                lam_i[:, i] += lamB_cfl[:, f1 - 1]
        d["dt"] = min(min(tp.CFL * tp.dt0, tp.T - t), float((tp.CFL * 0.5 * wJ / lam_i).min()))
    # ---- high order
    beta = u_tilde[..., 0] / (2 * p1(g, u_tilde))
    rl, bl = LOG(u_tilde[..., 0]), LOG(beta)
    uu = u_tilde[..., 1] / u_tilde[..., 0]
    lamf = ws1(g, utf)
    LFc = (0.5 * nn * np.maximum(lamf, lamf.reshape(K * 2)[mapP].reshape(K, 2))).reshape(K * 2)
    uPh = utf.reshape(K * 2, 3)[mapP].copy()
    if len(mapI):
        LFc[mapI] = 0.0
        uPh[mapI] = Ival
    if len(mapO):
        LFc[mapO] = 0.0
        uPh[mapO] = Uq[mapO // 2, fq2q[mapO % 2]]
    LFc, uPh = LFc.reshape(K, 2), uPh.reshape(K, 2, 3)
    Sh = rxJ[:, :, None] * np.asarray(ops.Srsh_db[0])[None]
    if vol_flux == T.VOLFLUX_CENTRAL:
        fh = flux1(g, u_tilde)
        Fij = 0.5 * (fh[:, :, None] + fh[:, None, :])
    else:
        L = lambda a: a[:, :, None]
        R = lambda a: a[:, None, :]
        Fij = fS1(g, L(u_tilde[..., 0]), L(uu), L(beta), L(rl), L(bl), R(u_tilde[..., 0]), R(uu), R(beta), R(rl), R(bl))
    QF1 = np.einsum("kij,kijc->kic", Sh, Fij)
    if high_flux == T.SURFFLUX_CHANDRASHEKAR_PROJECTED:
        bP = uPh[..., 0] / (2 * p1(g, uPh))
        fstar = fS1(g, utf[..., 0], utf[..., 1] / utf[..., 0], beta[:, Nq:], rl[:, Nq:], bl[:, Nq:],
                    uPh[..., 0], uPh[..., 1] / uPh[..., 0], bP, LOG(uPh[..., 0]), LOG(bP))
    else:
        fstar = 0.5 * (flux1(g, utf) + flux1(g, uPh))
    BF_H = Bx[..., None] * fstar - LFc[..., None] * (uPh - utf)
    if (theta_local == 1.0).all():
        proj = np.einsum("qh,khc->kqc", ops.MinvVhT, QF1) + np.einsum("qf,kfc->kqc", ops.MinvVfT, BF_H)
    else:       # the limited face matrix, flux_differencing.jl:288-319
        proj = (QF1[:, :Nq] + np.einsum("kfq,kfc->kqc", Vf_new, QF1[:, Nq:]) + np.einsum("kfq,kfc->kqc", Vf_new, BF_H)) / np.asarray(ops.wq)[None, :, None]
    rhsH = -proj / np.asarray(geom.Jq, dtype=float)[..., None]
    d["rhsH"], d["theta_local"] = rhsH, theta_local
    # ---- limiter
    lim = param.rhs_limiter
    subcell = lim.code == T.LIMITER_SUBCELL
    bcode = lim.bound.code if subcell else T.BOUND_POSITIVITY
    cell = bcode in (T.BOUND_POS_CELL_ENTROPY, T.BOUND_POS_RELAXED_CELL_ENTROPY, T.BOUND_TVD_CELL_ENTROPY, T.BOUND_TVD_RELAXED_CELL_ENTROPY)
    relaxed_cell = bcode in (T.BOUND_POS_RELAXED_CELL_ENTROPY, T.BOUND_TVD_RELAXED_CELL_ENTROPY)
    tvd = bcode >= T.BOUND_TVD
    minent = bcode in (T.BOUND_POS_MIN_ENTROPY, T.BOUND_POS_RELAXED_MIN_ENTROPY, T.BOUND_TVD_MIN_ENTROPY, T.BOUND_TVD_RELAXED_MIN_ENTROPY)
    relaxed = bcode in (T.BOUND_POS_RELAXED_MIN_ENTROPY, T.BOUND_TVD_RELAXED_MIN_ENTROPY)
    hen = lim.shockcapture.code == T.SHOCKCAPTURE_HENNEMANN
    blend = np.ones(K)
    if hen or relaxed or relaxed_cell:                                # initialize_smoothness_indicator!(::Dim1) shock_capture.jl:14-45
        modal = np.einsum("mq,kq->km", ops.VDM_inv, Uq[..., 0] * p1(g, Uq))
        e = modal ** 2
        tot = np.zeros(K)
        for m in range(e.shape[1]):
            tot = tot + e[:, m]
        sigma = np.maximum(e[:, param.N] / tot, e[:, param.N - 1] / tot)
    if hen:
        TN = lim.shockcapture.a * 10 ** (-lim.shockcapture.c * (param.N + 1) ** 0.25)
        blend = np.maximum(np.minimum(1.0 - 1.0 / (1.0 + EXP(-math.log((1 - 0.0001) / 0.0001) / TN * (sigma - TN))), 1.0), 0.5)
    uL = Uq + dt * rhsL
    if not subcell:
        L = lim1(ZEROTOL, uL, dt * (rhsH - rhsL), zeta * uL[..., 0], zeta * rhoe1(uL)).min(axis=1)
        l = np.minimum(L, blend)
        d["L"], d["rhsU"] = L, (1 - l)[:, None, None] * rhsL + l[:, None, None] * rhsH
        return d
    if minent or tvd:                                                 # low_order_stencil(::Dim1) limiter_utils.jl:212-220
        mp = np.asarray(bc.mapP).reshape(K, 2) - 1
        kN, qN = np.empty((K, Nq, 2), dtype=np.int64), np.empty((K, Nq, 2), dtype=np.int64)
        for q in range(Nq):
            for s_, (inside, qin) in enumerate(((q - 1 >= 0, q - 1), (q + 1 <= Nq - 1, q + 1))):
                if inside:
                    kN[:, q, s_], qN[:, q, s_] = np.arange(K), qin
                else:
                    P = mp[:, ops.q2fq[q][0] - 1]
                    kN[:, q, s_], qN[:, q, s_] = P // 2, fq2q[P % 2]
    Lphi = None
    if minent:
        sm = rhoe1(Uq) * POW(Uq[..., 0], -g)
        lb = np.minimum(sm, sm[kN, qN].min(axis=2))
        epsk = smooth_factor(param, sigma) if relaxed else np.ones(K)
        Lphi = epsk[:, None] * lb + (1 - epsk[:, None]) * smin
    Lrho, Urho = zeta * uL[..., 0], None
    if tvd:
        rhoL = Uq[..., 0] + dt * rhsL[..., 0]
        nb = rhoL[kN, qN]
        Lrho, Urho = np.minimum(rhoL, nb.min(axis=2)), np.maximum(rhoL, nb.max(axis=2))
    Lrhoe = zeta * rhoe1(uL)

    def coef(Pv):
        with np.errstate(divide="ignore", invalid="ignore"):
            l = np.where(uL[..., 0] + Pv[..., 0] < Lrho, np.maximum((Lrho - uL[..., 0]) / Pv[..., 0], 0.0), 1.0)
            if Urho is not None:
                l = np.where(uL[..., 0] + Pv[..., 0] > Urho, np.minimum(l, np.maximum((Urho - uL[..., 0]) / Pv[..., 0], 0.0)), l)
        l = np.minimum(np.minimum(l, quad1(ZEROTOL, uL, Pv, Lrhoe)), 1.0)
        if minent:
            def f(x):
                with np.errstate(all="ignore"):
                    w = uL + x[..., None] * Pv
                    return rhoe1(w) * POW(w[..., 0], -g) >= Lphi - POSTOL
            top = f(l)
            xv, xi = np.zeros_like(l), l.copy()
            for _ in range(21):
                xn = 0.5 * (xv + xi)
                good = f(xn)
                xv, xi = np.where(good, xn, xv), np.where(good, xi, xn)
            l = np.where(top, l, xv)
        return l
    fH, fL = np.zeros((K, Nq + 1, 3)), np.zeros((K, Nq + 1, 3))       # accumulate_f_bar!(::Dim1) :144-161
    fH[:, 0], fL[:, 0] = BF_H[:, 0], BF_L[:, 0]
    for i in range(1, Nq + 1):
        fH[:, i] = fH[:, i - 1] + wJ[:, i - 1, None] * rhsH[:, i - 1]
        fL[:, i] = fL[:, i - 1] + wJ[:, i - 1, None] * rhsL[:, i - 1]
    df = fH - fL
    Ll = np.ones((K, Nq + 1))                                         # subcell_bound_limiter!(::Dim1) :208-246
    Ll[:, :Nq] = np.minimum(Ll[:, :Nq], coef(-2 * dt * df[:, :Nq] / wJ[..., None]))
    Ll[:, 1:] = np.minimum(Ll[:, 1:], coef(2 * dt * df[:, 1:] / wJ[..., None]))
    Ll = np.minimum(Ll, blend[:, None])
    if cell:                                                          # enforce_ES_subcell!(::Dim1) :468-506, 567-628
        epsk = smooth_factor(param, sigma) if relaxed_cell else np.zeros(K)
        psif = (g - 1.0) * Uq[:, fq2q][..., 1]
        for k in range(K):
            sB = 0.0
            for f in range(2):
                sB += Bx[k, f] * psif[k, f]
            dv = {e: vq[k, e] - vq[k, e + 1] for e in range(Nq - 1)}                     # interior face e + 1 between nodes e, e + 1
            dvdf = {e: float(np.sum(dv[e] * (fH[k, e + 1] - fL[k, e + 1]))) for e in dv}
            sdvfL = 0.0
            for e in dv:
                sdvfL += float(np.sum(dv[e] * fL[k, e + 1]))
            spos = 0.0
            for e in dv:
                spos += Ll[k, e + 1] * dvdf[e]
            budget = sB - sdvfL
            rhs_ = (1 - lim.bound.beta * epsk[k]) * budget if relaxed_cell else budget
            tol = max(0.0, sdvfL - sB)
            if spos - rhs_ > tol:
                lhs, taken = spos, []
                for e in sorted(dvdf, key=lambda e_: (dvdf[e_], e_), reverse=True):
                    if not lhs > rhs_ + tol or dvdf[e] < ZEROTOL:
                        break
                    lhs -= Ll[k, e + 1] * dvdf[e]
                    taken.append(e)
                for m, e in enumerate(taken):
                    l_new = max((rhs_ + tol - lhs) / dvdf[e], 0.0) if m == len(taken) - 1 else 0.0
                    Ll[k, e + 1] = min(Ll[k, e + 1], l_new)
    first, last = Ll[:, 0].copy(), Ll[:, Nq].copy()                   # symmetrize_limiting_parameters!(::Dim1) :405-416
    l = np.minimum(first, np.roll(last, 1))
    Ll[:, 0] = l
    Ll[:, Nq] = np.roll(l, -1)
    flim = Ll[..., None] * fH + (1 - Ll[..., None]) * fL
    d["rhsU"] = (flim[:, 1:] - flim[:, :-1]) / wJ[..., None]
    d["Ll"] = Ll
    return d


def dense_ssp33_step(param, dd, bc, Uq, t, smin=None):
    """One iteration of the loop in SSP33! (timestepping/SSPRK33.jl:28-40) on top of the dense rhs!: dt capped by CFL dt0 and
    T - t, stage 1 returns the CFL dt but its limiter has seen the cap (rhs.jl:46,52), three stages with the SSP combinations.
    2D, LimitedDG.  -> (U_new, dt)"""
    tp = param.timestepping_param
    dt = min(tp.CFL * tp.dt0, tp.T - t)
    resW = Uq
    d = dense_limited_rhs(param, dd, bc, Uq, t, dt, 1, smin=smin)
    dt = d["dt"]
    U = resW + dt * d["rhsU"]
    d = dense_limited_rhs(param, dd, bc, U, t, dt, 2, smin=smin)
    resZ = U + dt * d["rhsU"]
    U = 3 / 4 * resW + 1 / 4 * resZ
    d = dense_limited_rhs(param, dd, bc, U, t, dt, 3, smin=smin)
    resZ = U + dt * d["rhsU"]
    return 1 / 3 * resW + 2 / 3 * resZ, dt
