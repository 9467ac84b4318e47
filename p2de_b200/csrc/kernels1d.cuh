// kernels1d.cuh — the per-stage hot path on 1D uniform line meshes.
//
// 1D problems of the reference (Sod / Shu-Osher / Leblanc shock tubes) have a few hundred to a few
// thousand elements, so this path is written for generality, not speed: one thread per element,
// operator-driven (dense Srh_db / S0 / Vf / M^-1 Vh^T exactly as passed by the caller), which
// covers Lobatto AND Gauss collocation.  Same two-kernel structure as 2D:
//   stage1d_kernel  = entropy projection (rhs.jl:59-133) + low-order RHS and CFL
//                     (low_order_graph_viscosity.jl:4-243) + flux differencing
//                     (flux_differencing.jl:4-361) + element-local limiter part
//                     (zhangshu.jl:4-45, subcell.jl:144-160,208-246)
//   update1d_kernel = symmetrisation (subcell.jl:405-416, hard-coded periodic neighbours k-1/k+1
//                     like the reference), limited reassembly (subcell.jl:826-892), SSP combine.
// Reference-order arithmetic (1D formulas of compressible_Navier_Stokes.jl:1-218).
// HEAD is broken in 1D (Bx argument order rhs_utils.jl:18, shadowed dim/bound subcell.jl:230-231); the
// evident intent is implemented.
#pragma once
#include "kernels2d.cuh"

namespace p2de {

struct Cons1 { double rho, m, E; };

template <int N1D>
struct Tables1D {
  static constexpr int Nq = N1D, Nh = N1D + 2;
  double Srh[Nh][Nh];     // Srh_db[i][j] (math indexed)
  double S0[Nq][Nq];      // low-order S0
  double Br[2];
  double Vf[2][Nq];
  double Vf_low[2][Nq];   // nearest-node gather (init.jl:181-213); blended with Vf by NodewiseScaledExtrapolation
  double MinvVhT[Nq][Nh];
  double MinvVfT[Nq][2];
  double wq[Nq];
  int fq2q[2];
  int vf_is_gather;       // Vf rows are exact unit vectors (Lobatto)
};

struct Args1D {
  const double *Uq;                  // [K][Nq][3]
  double *rhsL, *dF, *lpre;          // subcell scratch: [K][Nq][3], [K][Nq+1][3], [K][Nq+1]
  double *rhsU;                      // other modes / outputs
  double *Lout, *rhsH_diag, *rhsL_diag;
  unsigned long long *dt_bits;
  const double *dt_dev;
  double dt_host;
  int use_dt_dev, nstage;
  long long K;
  const int *mapP32;                 // [K][2] 0-based linear index into [2][K]
  const int *bcflag;                 // [K][2]: 0 none, >0 inflow (index+1 into Ival), -1 outflow
  const double *Ival;                // [nI][3]
  double gamma, ZEROTOL, POSTOL, zeta, CFL, Jq, rxJ, blend;
  int mode, vol_flux, surf_low, surf_high, roundtrip;
  // remaining subcell bounds and shock capturing (Solver.jl:47-74), as in the generic 2D kernel
  int tvd;                           // TVD*Bound: rho in [min, max] of the low-order update over the low-order stencil
  const double *rhsLpre;             // [K][Nq][3] low-order rhs of ALL elements (MODE_LOW pre-pass): the stencil crosses faces
  int entropy_bound;                 // 0 none, 1 PositivityAndMinEntropyBound, 2 ...RelaxedMinEntropyBound (and the TVD variants)
  int cell_entropy;                  // 0 none, 1 *CellEntropyBound, 2 *RelaxedCellEntropyBound(beta)
  int hennemann, N;
  double hen_a, hen_c, bound_beta;
  const double *VDM_inv;             // [Np, Nq] column-major
  const double *smin_dev;            // global minimum of s_modified at t0
  // NodewiseScaledExtrapolation (filter.jl:6-130): theta per face node, this stage's slots [K][2] / [K] (null: not stored)
  int nodewise, gauss;
  double eta;
  double *theta_local, *theta;
};

struct Upd1D {
  const double *rhsL, *dF, *lpre, *rhsU_in;
  double *Llocal_out, *rhsU_out;
  const double *Uq_in, *resW;
  double *Uq_out;
  double a, b;
  const double *dt_dev;
  double dt_host;
  int use_dt_dev, mode;
  long long K;
  double Jq;
};

// ---- 1D physics, compressible_Navier_Stokes.jl -------------------------------------------------
P2DE_DEV double p1(double gm1, const Cons1 &U) { return gm1 * (U.E - 0.5 * (U.m * U.m) / U.rho); }          // :18-22
P2DE_DEV double rhoe1(const Cons1 &U) { return U.E - 0.5 * U.m * U.m / U.rho; }                           // :70-73
P2DE_DEV double ws1(double gamma, double gm1, const Cons1 &U) {                                          // :48-52
  return fabs(U.m / U.rho) + sqrt(gamma * p1(gm1, U) / U.rho);
}
P2DE_DEV void flux1(double gm1, const Cons1 &U, double f[3]) {                                           // :165-173
  double p = p1(gm1, U), u = U.m / U.rho;
  f[0] = U.m; f[1] = U.m * u + p; f[2] = u * (U.E + p);
}
P2DE_DEV void v_of_u1(double gamma, double gm1, const Cons1 &U, double V[3]) {                           // :113-122
  double p = p1(gm1, U), s = log(p / pow(U.rho, gamma));
  V[0] = (gamma + 1 - s) - gm1 * U.E / p; V[1] = U.m * gm1 / p; V[2] = -U.rho * gm1 / p;
}
P2DE_DEV Cons1 u_of_v1(double gamma, double gm1, const double V[3]) {                                    // :87-104,146-153
  double s = gamma - V[0] + (V[1] * V[1]) / (2 * V[2]);
  double rhoeV = pow(gm1 / pow(-V[2], gamma), 1 / gm1) * exp(-s / gm1);
  Cons1 U; U.rho = -rhoeV * V[2]; U.m = rhoeV * V[1]; U.E = rhoeV * (1 - (V[1] * V[1]) / (2 * V[2]));
  return U;
}
struct Prim1 { double rho, u, beta, rholog, betalog; };
P2DE_DEV Prim1 prim1(double gm1, const Cons1 &U) {
  Prim1 q; q.rho = U.rho; q.u = U.m / U.rho; q.beta = U.rho / (2 * p1(gm1, U));
  q.rholog = log(U.rho); q.betalog = log(q.beta);
  return q;
}
P2DE_DEV void fS1(double gm1, const Prim1 &L, const Prim1 &R, double F[3]) {                             // :196-218
  double rholog = logmean(L.rho, R.rho, L.rholog, R.rholog);
  double betalog = logmean(L.beta, R.beta, L.betalog, R.betalog);
  double rhoavg = 0.5 * (L.rho + R.rho), uavg = 0.5 * (L.u + R.u), unorm = L.u * R.u;
  double pa = rhoavg / (L.beta + R.beta);
  double f4aux = rholog / (2 * gm1 * betalog) + pa + 0.5 * rholog * unorm;
  double F1 = rholog * uavg;
  F[0] = F1; F[1] = F1 * uavg + pa; F[2] = f4aux * uavg;
}
P2DE_DEV double s_modified1(double gamma, const Cons1 &U) { return rhoe1(U) * pow(U.rho, -gamma); }          // :80-85
// limiting_param_bound_rho_rhoe with a finite upper density bound (limiter_utils.jl:26-40; Urhoe = Inf gives 1)
P2DE_DEV double limiting_param_rho_bounds1(double ZEROTOL, const Cons1 &U, const double Pv[3], double Lrho, double Lrhoe, double Urho) {
  double l = 1.0;
  if (U.rho + Pv[0] < Lrho) l = jl_max((Lrho - U.rho) / Pv[0], 0.0);
  if (U.rho + Pv[0] > Urho) l = jl_min(l, jl_max((Urho - U.rho) / Pv[0], 0.0));
  double a = Pv[0] * Pv[2] - 1.0 / 2.0 * (Pv[1] * Pv[1]);
  double b = U.E * Pv[0] + U.rho * Pv[2] - U.m * Pv[1] - Pv[0] * Lrhoe;
  double c = U.E * U.rho - 1.0 / 2.0 * (U.m * U.m) - U.rho * Lrhoe;
  l = jl_min(l, rhoe_quadratic_roots(ZEROTOL, a, b, c));
  return jl_min(l, 1.0);
}
// limiting_param_bound_phi (limiter_utils.jl:42-50) with the reference's bisection (nonlinear_solvers.jl: 20 halvings,
// returns the last lower end that satisfied the predicate)
P2DE_DEV double limiting_param_phi1(double gamma, double POSTOL, const Cons1 &U, const double Pv[3], double Lphi, double lpos) {
  auto f = [&](double l) {
    Cons1 W; W.rho = U.rho + l * Pv[0]; W.m = U.m + l * Pv[1]; W.E = U.E + l * Pv[2];
    return s_modified1(gamma, W) >= Lphi - POSTOL;
  };
  if (f(lpos)) return lpos;
  double x_valid = 0.0, x_invalid = lpos;
  for (int iter = 0; iter <= 20; ++iter) {
    double x_new = 0.5 * (x_valid + x_invalid);
    if (f(x_new)) x_valid = x_new; else x_invalid = x_new;
  }
  return x_valid;
}
// limiter_utils.jl:26-90 (Dim1 coefficients :78-83), positivity bounds (Urho = Urhoe = Inf)
P2DE_DEV double limiting_param_pos1(double ZEROTOL, const Cons1 &U, const double Pv[3], double Lrho, double Lrhoe) {
  double l = 1.0;
  if (U.rho + Pv[0] < Lrho) l = jl_max((Lrho - U.rho) / Pv[0], 0.0);
  double a = Pv[0] * Pv[2] - 1.0 / 2.0 * (Pv[1] * Pv[1]);
  double b = U.E * Pv[0] + U.rho * Pv[2] - U.m * Pv[1] - Pv[0] * Lrhoe;
  double c = U.E * U.rho - 1.0 / 2.0 * (U.m * U.m) - U.rho * Lrhoe;
  l = jl_min(l, rhoe_quadratic_roots(ZEROTOL, a, b, c));
  return jl_min(l, 1.0);
}
P2DE_DEV double find_alpha1(double POSTOL, const Cons1 &ui, const Cons1 &ut) {   // low_order_graph_viscosity.jl:299-327
  if (!P2DE_FIND_ALPHA_BISECT) {   // closed form of what the bisection converges to (kernels2d.cuh: find_alpha_closed)
    const double dr = ut.rho - ui.rho, dm = ut.m - ui.m, dE = ut.E - ui.E;
    return find_alpha_closed<void>(POSTOL, ui.rho, ui.E, ui.m * ui.m, dr, dE, ui.m * dm, dm * dm);
  }
  double alphaL = 0.0, alphaR = 1.0;
  Cons1 s;
  auto sub = [&](double al) { s.rho = al * ui.rho - ut.rho; s.m = al * ui.m - ut.m; s.E = al * ui.E - ut.E; };
  sub(alphaR);
  while (!(s.rho > POSTOL && rhoe1(s) > POSTOL) && alphaR < 1e300) { alphaR = 2 * alphaR; sub(alphaR); }
  for (int it = 0; it < 50; ++it) {
    double alphaM = (alphaL + alphaR) / 2;
    sub(alphaM);
    if (s.rho > POSTOL && rhoe1(s) > POSTOL) alphaR = alphaM; else alphaL = alphaM;
  }
  return alphaR;
}

P2DE_DEV Cons1 load1(const double *p) { Cons1 U; U.rho = p[0]; U.m = p[1]; U.E = p[2]; return U; }

// entropy-projected state of element `Uel` at face f (rhs.jl:84-94): u(sum_j (th Vf + (1 - th) Vf_low)[f][j] v_j); th = 1
// without NodewiseScaledExtrapolation
template <int N1D>
P2DE_DEV Cons1 project_face(const Tables1D<N1D> &T, double gamma, double gm1, int roundtrip, const Cons1 Uel[N1D], int f,
                            int nodewise = 0, double th = 1.0) {
  if (T.vf_is_gather && !roundtrip) return Uel[T.fq2q[f]];
  double acc[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < N1D; ++j) {
    double V[3];
    v_of_u1(gamma, gm1, Uel[j], V);
    double w = nodewise ? th * T.Vf[f][j] + (1 - th) * T.Vf_low[f][j] : T.Vf[f][j];
    acc[0] += w * V[0]; acc[1] += w * V[1]; acc[2] += w * V[2];
  }
  return u_of_v1(gamma, gm1, acc);
}

// compute_entropyproj_limiting_param!(::GaussCollocation) for one face node (filter.jl:6-15,26-58): the largest theta in
// [0, 1] (1, or the valid end of 21 bisection steps, nonlinear_solvers.jl:3-20) whose blended extrapolation of the entropy
// variables keeps v3, rho and rho e inside the bounds relative to the plain extrapolation (check_bound_on_face_node :84-98,
// update_limited_entropyproj_vars_on_face_node! :110-130)
template <int N1D>
P2DE_DEV double theta_face1(const Tables1D<N1D> &T, double gamma, double gm1, double POSTOL, double zeta, double eta,
                            const Cons1 Uel[N1D], int f) {
  double V[N1D][3], Uf[3] = {0.0, 0.0, 0.0}, VUf[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < N1D; ++j) {
    v_of_u1(gamma, gm1, Uel[j], V[j]);
    const double w = T.Vf[f][j];
    Uf[0] += w * Uel[j].rho; Uf[1] += w * Uel[j].m; Uf[2] += w * Uel[j].E;
    VUf[0] += w * V[j][0]; VUf[1] += w * V[j][1]; VUf[2] += w * V[j][2];
  }
  Cons1 Ufc; Ufc.rho = Uf[0]; Ufc.m = Uf[1]; Ufc.E = Uf[2];
  const double rhoef = rhoe1(Ufc);
  auto ok = [&](double th) {
    double vt[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < N1D; ++j) {
      const double w = th * T.Vf[f][j] + (1 - th) * T.Vf_low[f][j];
      vt[0] += w * V[j][0]; vt[1] += w * V[j][1]; vt[2] += w * V[j][2];
    }
    if (!(vt[2] < -POSTOL)) return false;
    const Cons1 ut = u_of_v1(gamma, gm1, vt);
    const double rhoe = rhoe1(ut);
    return vt[2] < jl_min(zeta * VUf[2], -POSTOL) && ut.rho > jl_max((1 - eta) * Uf[0], POSTOL) && ut.rho < (1 + eta) * Uf[0] &&
           rhoe > jl_max((1 - eta) * rhoef, POSTOL) && rhoe < (1 + eta) * rhoef;
  };
  if (ok(1.0)) return 1.0;
  double x_valid = 0.0, x_invalid = 1.0;
  for (int it = 0; it <= 20; ++it) {
    const double x_new = 0.5 * (x_valid + x_invalid);
    if (ok(x_new)) x_valid = x_new; else x_invalid = x_new;
  }
  return x_valid;
}

template <int N1D>
__global__ void __launch_bounds__(64)
stage1d_kernel(const Args1D A, const Tables1D<N1D> T) {
  constexpr int Nq = N1D, Nh = N1D + 2;
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const double gamma = A.gamma, gm1 = A.gamma - 1.0;
  const bool do_low = A.mode != MODE_HIGH, do_high = A.mode != MODE_LOW;
  double dtloc = INFINITY;
  if (k < A.K) {
    Cons1 U[Nq];
    for (int i = 0; i < Nq; ++i) U[i] = load1(A.Uq + (k * Nq + i) * 3);
    // neighbour elements through mapP (self on non-periodic boundaries)
    Cons1 ut[2], utP[2], UnodeP[2], uP_L[2], uP_H[2];
    int bc[2];
    const double *ival[2];
    long long kPs[2]; int nodeP[2];      // low-order stencil across the faces (limiter_utils.jl:184-207): partner element / node
    // NodewiseScaledExtrapolation: theta of the own faces and (recomputed from the neighbour's nodes with the same arithmetic,
    // hence the same bits as the neighbour's own value) of the partner faces; Lobatto: 1 (filter.jl:18-20)
    const bool nw = A.nodewise && A.gauss;
    double th[2] = {1.0, 1.0};
    for (int f = 0; f < 2; ++f) {
      int m = A.mapP32[k * 2 + f];
      long long kP = m / 2; int fP = m % 2;
      kPs[f] = kP; nodeP[f] = T.fq2q[fP];
      Cons1 Unb[Nq];
      for (int i = 0; i < Nq; ++i) Unb[i] = load1(A.Uq + (kP * Nq + i) * 3);
      double thP = 1.0;
      if (nw) {
        th[f] = theta_face1<N1D>(T, gamma, gm1, A.POSTOL, A.zeta, A.eta, U, f);
        thP = theta_face1<N1D>(T, gamma, gm1, A.POSTOL, A.zeta, A.eta, Unb, fP);
      }
      ut[f] = project_face<N1D>(T, gamma, gm1, A.roundtrip, U, f, nw, th[f]);
      utP[f] = project_face<N1D>(T, gamma, gm1, A.roundtrip, Unb, fP, nw, thP);
      UnodeP[f] = Unb[T.fq2q[fP]];
      int fl = A.bcflag ? A.bcflag[k * 2 + f] : 0;
      bc[f] = fl > 0 ? 1 : (fl < 0 ? 2 : 0);
      ival[f] = fl > 0 ? A.Ival + 3ll * (fl - 1) : nullptr;
    }
    if (A.nodewise && A.theta_local) {
      A.theta_local[k * 2 + 0] = th[0]; A.theta_local[k * 2 + 1] = th[1];
      if (A.gauss && A.theta) A.theta[k] = ((0.0 + th[0]) + th[1]) / 2;   // filter.jl:57 (the Lobatto method never writes theta)
    }
    double rL[Nq][3], rH[Nq][3], BFL[2][3], BFH[2][3];
    for (int i = 0; i < Nq; ++i) for (int c = 0; c < 3; ++c) { rL[i][c] = 0.0; rH[i][c] = 0.0; }
    for (int f = 0; f < 2; ++f) for (int c = 0; c < 3; ++c) { BFL[f][c] = 0.0; BFH[f][c] = 0.0; }

    if (do_low) {
      // ---- low_order_graph_viscosity.jl:43-220
      const bool proj = A.surf_low == P2DE_SURFFLUX_LF_PROJECTED;
      double fl[Nq][3], Q0[Nq][3], lam[Nq][Nq];
      for (int i = 0; i < Nq; ++i) { flux1(gm1, U[i], fl[i]); for (int c = 0; c < 3; ++c) Q0[i][c] = 0.0; for (int j = 0; j < Nq; ++j) lam[i][j] = 0.0; }
      for (int j = 0; j < Nq; ++j)
        for (int i = j + 1; i < Nq; ++i) {
          if (T.S0[i][j] == 0.0) continue;
          double Sv = A.rxJ * T.S0[i][j], nn = fabs(Sv);
          double l = nn * jl_max(ws1(gamma, gm1, U[i]), ws1(gamma, gm1, U[j]));
          lam[i][j] = l; lam[j][i] = l;
          const double ui[3] = {U[i].rho, U[i].m, U[i].E}, uj[3] = {U[j].rho, U[j].m, U[j].E};
          for (int c = 0; c < 3; ++c) {
            double SF = 2.0 * Sv * (0.5 * (fl[i][c] + fl[j][c])) - l * (uj[c] - ui[c]);
            Q0[i][c] += SF; Q0[j][c] += -SF;
          }
        }
      for (int i = 0; i < Nq; ++i) for (int c = 0; c < 3; ++c) rL[i][c] = 0.0 - Q0[i][c];
      double lamB[2], lamFace[2];
      for (int f = 0; f < 2; ++f) {
        double B = T.Br[f] * A.rxJ, nn = fabs(B);
        Cons1 Uf = proj ? ut[f] : U[T.fq2q[f]];
        Cons1 UfP = proj ? utP[f] : UnodeP[f];
        double wsM = ws1(gamma, gm1, Uf), wsP = ws1(gamma, gm1, UfP);
        lamB[f] = 0.5 * nn * jl_max(wsM, wsP);
        Cons1 uP = UfP;
        if (bc[f] == 1) uP = load1(ival[f]); else if (bc[f] == 2) uP = U[T.fq2q[f]];
        uP_L[f] = uP;
        double fM[3], fP[3];
        flux1(gm1, Uf, fM); flux1(gm1, uP, fP);
        const double uf[3] = {Uf.rho, Uf.m, Uf.E}, up[3] = {uP.rho, uP.m, uP.E};
        for (int c = 0; c < 3; ++c) {
          BFL[f][c] = B * (0.5 * (fM[c] + fP[c])) - lamB[f] * (up[c] - uf[c]);
          rL[T.fq2q[f]][c] -= BFL[f][c];
        }
        lamFace[f] = lamB[f];
        if (proj && A.nstage == 1) lamFace[f] = find_alpha1(A.POSTOL, U[T.fq2q[f]], Uf) * lamB[f] + 0.5 * nn * wsM;
      }
      for (int i = 0; i < Nq; ++i) { double wJ = A.Jq * T.wq[i]; for (int c = 0; c < 3; ++c) rL[i][c] = rL[i][c] / wJ; }
      if (A.nstage == 1) {   // :222-291
        for (int i = 0; i < Nq; ++i) {
          double li = 0.0;
          for (int j = 0; j < Nq; ++j) li += lam[i][j];
          for (int f = 0; f < 2; ++f) if (T.fq2q[f] == i) li += lamFace[f];   // q2fq[i] (init.jl:215-220)
          dtloc = jl_min(dtloc, A.CFL * 0.5 * (A.Jq * T.wq[i]) / li);
        }
      }
    }

    if (do_high) {
      // ---- flux_differencing.jl:39-361 over the Nh = Nq + 2 hybridized points
      Cons1 uh[Nh];
      Prim1 q[Nh];
      for (int i = 0; i < Nq; ++i) uh[i] = U[i];
      uh[Nq] = ut[0]; uh[Nq + 1] = ut[1];
      for (int i = 0; i < Nh; ++i) q[i] = prim1(gm1, uh[i]);
      double QF[Nh][3];
      for (int i = 0; i < Nh; ++i) for (int c = 0; c < 3; ++c) QF[i][c] = 0.0;
      for (int j = 0; j < Nh; ++j)
        for (int i = j + 1; i < Nh; ++i) {
          double Sv = T.Srh[i][j];
          if (Sv == 0.0) continue;
          Sv = A.rxJ * Sv;
          double F[3];
          if (A.vol_flux == P2DE_VOLFLUX_CHANDRASHEKAR) fS1(gm1, q[i], q[j], F);
          else { double fi[3], fj[3]; flux1(gm1, uh[i], fi); flux1(gm1, uh[j], fj); for (int c = 0; c < 3; ++c) F[c] = 0.5 * (fi[c] + fj[c]); }
          for (int c = 0; c < 3; ++c) { double Sf = Sv * F[c]; QF[i][c] += Sf; QF[j][c] += -Sf; }
        }
      for (int f = 0; f < 2; ++f) {
        double B = T.Br[f] * A.rxJ, nn = fabs(B);
        double LFc = 0.5 * nn * jl_max(ws1(gamma, gm1, ut[f]), ws1(gamma, gm1, utP[f]));
        Cons1 uP = utP[f];
        if (bc[f] == 1) { uP = load1(ival[f]); LFc = 0.0; } else if (bc[f] == 2) { uP = U[T.fq2q[f]]; LFc = 0.0; }
        uP_H[f] = uP;
        double fs[3];
        if (A.surf_high == P2DE_SURFFLUX_CHANDRASHEKAR_PROJECTED) fS1(gm1, prim1(gm1, ut[f]), prim1(gm1, uP), fs);
        else { double fM[3], fP[3]; flux1(gm1, ut[f], fM); flux1(gm1, uP, fP); for (int c = 0; c < 3; ++c) fs[c] = 0.5 * (fM[c] + fP[c]); }
        const double uf[3] = {ut[f].rho, ut[f].m, ut[f].E}, up[3] = {uP.rho, uP.m, uP.E};
        for (int c = 0; c < 3; ++c) BFH[f][c] = B * fs[c] - LFc * (up[c] - uf[c]);
      }
      // assemble_rhs! (flux_differencing.jl:274-361); an element with some theta < 1 lifts with the limited face matrix
      // Vf_new = theta Vf + (1 - theta) Vf_low: M^-1 Vh^T -> (1/wq) [I Vf_new^T], M^-1 Vf^T -> (1/wq) Vf_new^T (:288-319)
      const bool limited = nw && jl_min(th[0], th[1]) < 1.0;
      for (int i = 0; i < Nq; ++i)
        for (int c = 0; c < 3; ++c) {
          double a = 0.0, b = 0.0;
          if (!limited) {
            for (int h = 0; h < Nh; ++h) a += T.MinvVhT[i][h] * QF[h][c];
            for (int f = 0; f < 2; ++f) b += T.MinvVfT[i][f] * BFH[f][c];
          } else {
            for (int h = 0; h < Nh; ++h) {
              const double vht = h < Nq ? (h == i ? 1.0 : 0.0)
                                        : th[h - Nq] * T.Vf[h - Nq][i] + (1 - th[h - Nq]) * T.Vf_low[h - Nq][i];
              a += ((1 / T.wq[i]) * vht) * QF[h][c];
            }
            for (int f = 0; f < 2; ++f) b += ((1 / T.wq[i]) * (th[f] * T.Vf[f][i] + (1 - th[f]) * T.Vf_low[f][i])) * BFH[f][c];
          }
          rH[i][c] = -(a + b) / A.Jq;
        }
    }
    (void)uP_L; (void)uP_H;

    const double dtl = A.use_dt_dev ? dt_read(A.dt_dev) : A.dt_host;
    // ---- smoothness indicator (shock_capture.jl:4-94), blending factor (:111-132), smoothness factor of the relaxed
    //      bounds (subcell.jl:927-956)
    double blend = A.blend, epsk = A.entropy_bound == 1 ? 1.0 : 0.0;
    if (A.hennemann || A.entropy_bound == 2 || A.cell_entropy == 2) {
      double ind[Nq], eN = 0.0, eNm1 = 0.0, etot = 0.0;
      for (int i = 0; i < Nq; ++i) ind[i] = U[i].rho * p1(gm1, U[i]);
      for (int m = 0; m < Nq; ++m) {
        double coef = 0.0;
        for (int j = 0; j < Nq; ++j) coef += A.VDM_inv[m + j * Nq] * ind[j];
        double e = coef * coef;
        if (m == A.N) eN += e;
        if (m == A.N - 1) eNm1 += e;
        etot += e;
      }
      const double sigma = jl_max(eN / etot, eNm1 / etot);
      if (A.hennemann) {
        const double TN = A.hen_a * pow(10.0, -A.hen_c * pow((double)(A.N + 1), 0.25));
        const double s_factor = log((1 - 0.0001) / 0.0001);
        const double al = 1.0 / (1.0 + exp(-s_factor / TN * (sigma - TN)));
        blend = jl_max(jl_min(1.0 - al, 1.0), 0.5);
      }
      if (A.entropy_bound == 2 || A.cell_entropy == 2) {
        const double kappa = 1.0, s0 = log10(pow((double)A.N, -4.0)), sk = log10(sigma);
        epsk = sk < s0 - kappa ? 0.0 : (sk > s0 + kappa ? 1.0 : 0.5 - 0.5 * sin(3.141592653589793 * (sk - s0) / (2 * kappa)));
      }
    }
    if (A.mode == MODE_SUBCELL) {
      // accumulate_f_bar! (subcell.jl:144-160) and subcell_bound_limiter!(::Dim1) (:208-246)
      double fH[3], fL[3], dF[Nq + 1][3];
      for (int c = 0; c < 3; ++c) { fH[c] = BFH[0][c]; fL[c] = BFL[0][c]; dF[0][c] = fH[c] - fL[c]; }
      for (int i = 1; i < Nq + 1; ++i)
        for (int c = 0; c < 3; ++c) {
          fH[c] = fH[c] + A.Jq * T.wq[i - 1] * rH[i - 1][c];
          fL[c] = fL[c] + A.Jq * T.wq[i - 1] * rL[i - 1][c];
          dF[i][c] = fH[c] - fL[c];
        }
      // ---- bounds over the low-order stencil {i-1, i+1} (neighbour element's face node across a face):
      //      minimum modified entropy (subcell.jl:37-55), TVD density interval (:86-110)
      double Lphi[Nq], Lrho_b[Nq], Urho_b[Nq];
      for (int i = 0; i < Nq; ++i) { Lphi[i] = 0.0; Lrho_b[i] = 0.0; Urho_b[i] = INFINITY; }
      if (A.entropy_bound) {
        double sm[Nq], smP[2];
        for (int i = 0; i < Nq; ++i) sm[i] = s_modified1(gamma, U[i]);
        for (int f = 0; f < 2; ++f) smP[f] = s_modified1(gamma, load1(A.Uq + (kPs[f] * Nq + nodeP[f]) * 3));
        const double smin = *A.smin_dev;
        for (int i = 0; i < Nq; ++i) {
          double lb = sm[i];
          lb = jl_min(lb, i > 0 ? sm[i - 1] : smP[0]);
          lb = jl_min(lb, i < Nq - 1 ? sm[i + 1] : smP[1]);
          Lphi[i] = epsk * lb + (1 - epsk) * smin;
        }
      }
      if (A.tvd) {
        double rhoL[Nq], rhoLP[2];
        for (int i = 0; i < Nq; ++i) rhoL[i] = U[i].rho + dtl * rL[i][0];
        for (int f = 0; f < 2; ++f) {
          const long long q = kPs[f] * Nq + nodeP[f];
          rhoLP[f] = A.Uq[q * 3] + dtl * A.rhsLpre[q * 3];
        }
        for (int i = 0; i < Nq; ++i) {
          const double a = i > 0 ? rhoL[i - 1] : rhoLP[0], b = i < Nq - 1 ? rhoL[i + 1] : rhoLP[1];
          Lrho_b[i] = jl_min(jl_min(rhoL[i], a), b);
          Urho_b[i] = jl_max(jl_max(rhoL[i], a), b);
        }
      }
      double lv[Nq + 1];
      for (int i = 0; i < Nq + 1; ++i) lv[i] = 1.0;
      for (int i = 0; i < Nq; ++i) {
        Cons1 uL; uL.rho = U[i].rho + dtl * rL[i][0]; uL.m = U[i].m + dtl * rL[i][1]; uL.E = U[i].E + dtl * rL[i][2];
        double wJ = T.wq[i] * A.Jq, Lrho = A.tvd ? Lrho_b[i] : A.zeta * uL.rho, Lrhoe = A.zeta * rhoe1(uL);
        double Pm[3], Pp[3];
        for (int c = 0; c < 3; ++c) { Pm[c] = -2 * dtl * dF[i][c] / wJ; Pp[c] = 2 * dtl * dF[i + 1][c] / wJ; }
        double lm = A.tvd ? limiting_param_rho_bounds1(A.ZEROTOL, uL, Pm, Lrho, Lrhoe, Urho_b[i]) : limiting_param_pos1(A.ZEROTOL, uL, Pm, Lrho, Lrhoe);
        double lp = A.tvd ? limiting_param_rho_bounds1(A.ZEROTOL, uL, Pp, Lrho, Lrhoe, Urho_b[i]) : limiting_param_pos1(A.ZEROTOL, uL, Pp, Lrho, Lrhoe);
        if (A.entropy_bound) {
          lm = limiting_param_phi1(gamma, A.POSTOL, uL, Pm, Lphi[i], lm);
          lp = limiting_param_phi1(gamma, A.POSTOL, uL, Pp, Lphi[i], lp);
        }
        lv[i] = jl_min(lv[i], lm);
        lv[i + 1] = jl_min(lv[i + 1], lp);
      }
      for (int i = 0; i < Nq + 1; ++i) lv[i] = jl_min(lv[i], blend);   // shock capturing (subcell.jl:243-245)
      if (A.cell_entropy) {
        // ---- enforce_ES_subcell!(::Dim1) (subcell.jl:458-707): entropy estimate of the element's interior subcell faces
        //      and the greedy fix in descending (dv.dF, index) order; no interface part in 1D
        double V[Nq][3], dv[Nq], fLr[3], fHr[3];
        for (int i = 0; i < Nq; ++i) v_of_u1(gamma, gm1, U[i], V[i]);
        double sBpsi = 0.0;
        for (int f = 0; f < 2; ++f) sBpsi += (T.Br[f] * A.rxJ) * (gm1 * U[T.fq2q[f]].m);   // psi_ufun :124-132 at the face's volume node
        double sdvfL = 0.0;
        for (int c = 0; c < 3; ++c) { fHr[c] = BFH[0][c]; fLr[c] = BFL[0][c]; }
        for (int si = 1; si < Nq; ++si) {
          for (int c = 0; c < 3; ++c) { fHr[c] = fHr[c] + A.Jq * T.wq[si - 1] * rH[si - 1][c]; fLr[c] = fLr[c] + A.Jq * T.wq[si - 1] * rL[si - 1][c]; }
          double a = 0.0, b = 0.0;
          for (int c = 0; c < 3; ++c) { const double dvc = V[si - 1][c] - V[si][c]; a += dvc * (fHr[c] - fLr[c]); b += dvc * fLr[c]; }
          dv[si - 1] = a; sdvfL += b;
        }
        const int n = Nq - 1;
        double sum_poslim = 0.0;
        for (int i = 0; i < n; ++i) sum_poslim += lv[i + 1] * dv[i];
        const double rhs_es = (A.cell_entropy == 1) ? sBpsi - sdvfL : (1 - A.bound_beta * epsk) * (sBpsi - sdvfL);
        const double tol = jl_max(0.0, sdvfL - sBpsi);
        if (sum_poslim - rhs_es > tol) {
          bool used[Nq];
          for (int i = 0; i < n; ++i) used[i] = false;
          double lhs = sum_poslim;
          int last = -1;
          while (lhs > rhs_es + tol) {
            int best = -1;   // the largest remaining (dv, index) tuple: sort!(..., rev=true) on tuples, Base.isless semantics
            for (int i = 0; i < n; ++i) {
              if (used[i]) continue;
              if (best < 0) { best = i; continue; }
              const double x = dv[i], y = dv[best];
              bool greater;
              if (x != x) greater = !(y != y) || i > best;
              else if (y != y) greater = false;
              else if (x == y) greater = (signbit(y) && !signbit(x)) || (signbit(x) == signbit(y) && i > best);
              else greater = x > y;
              if (greater) best = i;
            }
            if (best < 0 || dv[best] < A.ZEROTOL) break;
            lhs = lhs - lv[best + 1] * dv[best];
            used[best] = true; last = best;
          }
          for (int i = 0; i < n; ++i)
            if (used[i]) {
              const double l_new = (i == last) ? jl_max((rhs_es + tol - lhs) / dv[i], 0.0) : 0.0;
              lv[i + 1] = jl_min(lv[i + 1], l_new);
            }
        }
      }
      for (int i = 0; i < Nq + 1; ++i) {
        A.lpre[k * (Nq + 1) + i] = lv[i];
        for (int c = 0; c < 3; ++c) A.dF[(k * (Nq + 1) + i) * 3 + c] = dF[i][c];
      }
      for (int i = 0; i < Nq; ++i) for (int c = 0; c < 3; ++c) A.rhsL[(k * Nq + i) * 3 + c] = rL[i][c];
    } else {
      double l = 1.0;
      if (A.mode == MODE_ZHANGSHU) {   // zhangshu.jl:4-45
        for (int i = 0; i < Nq; ++i) {
          Cons1 uL; uL.rho = U[i].rho + dtl * rL[i][0]; uL.m = U[i].m + dtl * rL[i][1]; uL.E = U[i].E + dtl * rL[i][2];
          double Pv[3];
          for (int c = 0; c < 3; ++c) Pv[c] = dtl * (rH[i][c] - rL[i][c]);
          l = jl_min(l, limiting_param_pos1(A.ZEROTOL, uL, Pv, A.zeta * uL.rho, A.zeta * rhoe1(uL)));
        }
        A.Lout[k] = l;
        l = jl_min(l, blend);
      }
      for (int i = 0; i < Nq; ++i)
        for (int c = 0; c < 3; ++c)
          A.rhsU[(k * Nq + i) * 3 + c] = A.mode == MODE_ZHANGSHU ? (1 - l) * rL[i][c] + l * rH[i][c] : (A.mode == MODE_LOW ? rL[i][c] : rH[i][c]);
    }
    if (A.rhsL_diag && do_low) for (int i = 0; i < Nq; ++i) for (int c = 0; c < 3; ++c) A.rhsL_diag[(k * Nq + i) * 3 + c] = rL[i][c];
    if (A.rhsH_diag && do_high) for (int i = 0; i < Nq; ++i) for (int c = 0; c < 3; ++c) A.rhsH_diag[(k * Nq + i) * 3 + c] = rH[i][c];
  }
  if (do_low && A.nstage == 1) dt_publish(A.dt_bits, dtloc);
}

template <int N1D>
__global__ void __launch_bounds__(64)
update1d_kernel(const Upd1D A, const Tables1D<N1D> T) {
  constexpr int Nq = N1D;
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= A.K) return;
  const double dt = A.use_dt_dev ? dt_read(A.dt_dev) : A.dt_host;
  double r[Nq][3];
  if (A.mode == MODE_SUBCELL) {
    double lv[Nq + 1];
    for (int i = 0; i < Nq + 1; ++i) lv[i] = A.lpre[k * (Nq + 1) + i];
    // symmetrize_limiting_parameters!(::Dim1), subcell.jl:405-416: periodic neighbours, hard-coded
    const long long km = (k - 1 + A.K) % A.K, kp = (k + 1) % A.K;
    lv[0] = jl_min(lv[0], A.lpre[km * (Nq + 1) + Nq]);
    lv[Nq] = jl_min(lv[Nq], A.lpre[kp * (Nq + 1) + 0]);
    if (A.Llocal_out)   // reference shape [Nq + N1D, 1, K, Ns]; entries beyond Nq+1 are never limited (= 1)
      for (int i = 0; i < 2 * Nq; ++i) A.Llocal_out[k * (2 * Nq) + i] = i < Nq + 1 ? lv[i] : 1.0;
    for (int i = 0; i < Nq; ++i) {
      double wJ = T.wq[i] * A.Jq;
      for (int c = 0; c < 3; ++c)
        r[i][c] = A.rhsL[(k * Nq + i) * 3 + c] +
                  (lv[i + 1] * A.dF[(k * (Nq + 1) + i + 1) * 3 + c] - lv[i] * A.dF[(k * (Nq + 1) + i) * 3 + c]) / wJ;
    }
  } else {
    for (int i = 0; i < Nq; ++i) for (int c = 0; c < 3; ++c) r[i][c] = A.rhsU_in[(k * Nq + i) * 3 + c];
  }
  for (int i = 0; i < Nq; ++i)
    for (int c = 0; c < 3; ++c) {
      long long off = (k * Nq + i) * 3 + c;
      if (A.rhsU_out) A.rhsU_out[off] = r[i][c];
      if (A.Uq_out) {
        double uo = A.Uq_in[off];
        A.Uq_out[off] = (A.b == 1.0 && A.a == 0.0) ? uo + dt * r[i][c] : A.a * A.resW[off] + A.b * (uo + dt * r[i][c]);
      }
    }
}

__global__ void reduce1d_kernel(const double *U, const double *wq, int Nq, long long n_nodes, double J, int what, double *partial) {
  __shared__ double sh[256];
  double acc = what == P2DE_REDUCE_CONSERVATION ? 0.0 : INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_nodes; i += (long long)gridDim.x * blockDim.x) {
    Cons1 u = load1(U + i * 3);
    if (what == P2DE_REDUCE_CONSERVATION) acc += J * wq[i % Nq] * ((u.rho + u.m) + u.E);
    else if (what == P2DE_REDUCE_MIN_RHO) acc = fmin(acc, u.rho);
    else acc = fmin(acc, rhoe1(u));
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = what == P2DE_REDUCE_CONSERVATION ? sh[threadIdx.x] + sh[threadIdx.x + s] : fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

}  // namespace p2de
